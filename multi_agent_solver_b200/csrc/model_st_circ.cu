// model_st_circ.cu -- instantiates the batched iLQR kernels for the StCirc model (models.cuh).
#include "engine.cuh"

namespace mas_b200 {
BatchBase* make_batch_st_circ() { return new BatchImpl<StCirc>(); }
}  // namespace mas_b200
