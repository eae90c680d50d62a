// include/ark/AvatarRenderer.h -- the reference's header name (include/AvatarRenderer.h of sxyu/avatar), forwarding to the avatar_b200 facade so that
// a caller's `#include "AvatarRenderer.h"` resolves with -I<repo>/include/ark.
#pragma once
#include "../ark_b200/AvatarRenderer.h"
