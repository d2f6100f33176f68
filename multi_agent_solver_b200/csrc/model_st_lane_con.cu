// model_st_lane_con.cu -- instantiates the batched iLQR kernels for the StLaneCon model (models.cuh): lane
// following with one equality and one inequality path constraint, i.e. the augmented-Lagrangian code path.
#include "engine.cuh"

namespace mas_b200 {
BatchBase* make_batch_st_lane_con() { return new BatchImpl<StLaneCon>(); }
}  // namespace mas_b200
