// train.cu -- training-step kernels (placeholder until the kernels land)
#include "kernels.cuh"
namespace bfcnn {
int run_corrupt(bfcnn_handle*, const uint8_t*, float*, float*, int, int, int, uint64_t, uint64_t,
                const bfcnn_noise_cfg*, cudaStream_t) { set_error("corrupt: not built yet"); return BFCNN_ERR_UNSUPPORTED; }
int run_loss(bfcnn_handle*, const float*, const float*, int, int, int, const bfcnn_loss_cfg*, float*, cudaStream_t) { set_error("loss: not built yet"); return BFCNN_ERR_UNSUPPORTED; }
int run_train_step(bfcnn_handle*, const float*, const float*, int, int, int, const bfcnn_loss_cfg*, float*, float*, int, cudaStream_t) { set_error("train_step: not built yet"); return BFCNN_ERR_UNSUPPORTED; }
int run_adam_step(bfcnn_handle*, const float*, float, const bfcnn_adam_cfg*, int64_t, cudaStream_t) { set_error("adam: not built yet"); return BFCNN_ERR_UNSUPPORTED; }
}
