// model_st_lane.cu -- instantiates the batched iLQR kernels for the StLane model (models.cuh).
#include "engine.cuh"

namespace mas_b200 {
BatchBase* make_batch_st_lane() { return new BatchImpl<StLane>(); }
}  // namespace mas_b200
