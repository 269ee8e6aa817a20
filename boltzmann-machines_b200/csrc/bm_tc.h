// Tensor-core (tcgen05) path: declarations.  See bm_tc.cu.
#pragma once
#include "bm_internal.h"
