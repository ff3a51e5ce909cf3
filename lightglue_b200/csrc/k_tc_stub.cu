// placeholder until the tcgen05 path lands
#include "lg_handle.h"
int tc_pack_weights(LgHandle*, cudaStream_t) { return lg_set_error("tensor-core path not built"); }
void tc_free_weights(TcWeights*) {}
void tc_carve(size_t*, char*, size_t, int, const LgHandle*, TcBuffers*) {}
int tc_refresh_shadow(LgHandle*, const TcBuffers&, const float*, const SeqState&, cudaStream_t) { return lg_set_error("tc"); }
int tc_input_proj(LgHandle*, const TcBuffers&, const SeqState&, const float*, float*, cudaStream_t) { return lg_set_error("tc"); }
int tc_block(LgHandle*, const TcBuffers&, const SeqState&, int, int, float*, const float*, cudaStream_t) { return lg_set_error("tc"); }
int tc_final_proj(LgHandle*, const TcBuffers&, const SeqState&, float*, cudaStream_t) { return lg_set_error("tc"); }
