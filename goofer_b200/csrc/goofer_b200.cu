// goofer_b200.cu -- single translation unit of libgoofer_b200.so (the kernels share one __device__
// table object and a handful of inline device helpers, so they are compiled together).
#include "k_prep.cu"
#include "k_conv.cu"
#include "k_excite.cu"
#include "k_frame.cu"
#include "k_stage.cu"
#include "k_tail.cu"
#include "api.cu"
#include "k_fx.cu"
#include "host_api.cu"
#include "k_analyse.cu"
