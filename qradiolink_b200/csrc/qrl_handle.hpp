// qrl_handle.hpp -- the part every handle type of libqrl_b200.so shares (qrl_rx, qrl_tx, qrl_pfb, qrl_deframer, qrl_dfbb all derive
// from it as their only base), so that qrl_last_error(handle) reads `err` through a real base class instead of a layout pun.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

struct QrlHandleBase {
    std::string err;
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    long launches = 0;
    std::vector<void*> allocs;
};
