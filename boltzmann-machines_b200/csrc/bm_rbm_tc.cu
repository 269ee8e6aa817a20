#include "bm_internal.h"
namespace bm {
struct RbmBase;
RbmBase* make_rbm_tc(Ctx*, const bm_rbm_cfg&) {
    throw Error(BM_EUNSUPPORTED, "bf16 tensor-core engine not built yet");
}
}
