// include/ark/Calibration.h -- the reference's header name (include/Calibration.h of sxyu/avatar), forwarding to the avatar_b200 facade so that
// a caller's `#include "Calibration.h"` resolves with -I<repo>/include/ark.
#pragma once
#include "../ark_b200/AvatarOptimizer.h"
