// Single translation unit of libpotus_b200.so: device kernels + C-ABI host code.
#include "potus_kernel.cu"
#include "potus_stream.cu"
#include "potus_post.cu"
#include "potus_host.cu"
