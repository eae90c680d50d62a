// include/ark/Avatar.h -- the reference's header name (include/Avatar.h of sxyu/avatar), forwarding to the avatar_b200 facade so that
// a caller's `#include "Avatar.h"` resolves with -I<repo>/include/ark.
#pragma once
#include "../ark_b200/Avatar.h"
