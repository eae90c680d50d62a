// include/ark/RTree.h -- the reference's header name (include/RTree.h of sxyu/avatar), forwarding to the avatar_b200 facade so that
// a caller's `#include "RTree.h"` resolves with -I<repo>/include/ark.
#pragma once
#include "../ark_b200/RTree.h"
