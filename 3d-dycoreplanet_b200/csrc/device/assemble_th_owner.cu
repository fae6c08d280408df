// Row-owner ("written once") assembly of the classic NSE system / preconditioner matrices.
// Placeholder until the tile plan lands: selecting DCP_STRATEGY_OWNER fails loudly.
#include "dcp_internal.cuh"

struct OwnerPlan {
  int dummy;
};

int dcp_owner_plan_build(dcp_model*, bool, const dcp_model_desc*) {
  dcp_set_error("row-owner strategy is not built yet");
  return DCP_ERR_STATE;
}
void dcp_owner_plan_free(OwnerPlan* p) { delete p; }
int dcp_launch_th_owner(dcp_model*, const dcp_params&, bool) {
  dcp_set_error("row-owner strategy is not built yet");
  return DCP_ERR_STATE;
}
