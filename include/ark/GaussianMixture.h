// include/ark/GaussianMixture.h -- the reference's header name (include/GaussianMixture.h of sxyu/avatar), forwarding to the avatar_b200 facade so that
// a caller's `#include "GaussianMixture.h"` resolves with -I<repo>/include/ark.
#pragma once
#include "../ark_b200/Avatar.h"
