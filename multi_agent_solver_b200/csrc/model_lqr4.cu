// model_lqr4.cu -- instantiates the batched iLQR kernels for the Lqr4 model (models.cuh).
#include "engine.cuh"

namespace mas_b200 {
BatchBase* make_batch_lqr4() { return new BatchImpl<Lqr4>(); }
}  // namespace mas_b200
