// Host-side helpers shared by the translation units of libtrafficbots_b200.so.
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/trafficbots_b200.h"
#include "tb_device.cuh"
#include "tb_tc.cuh"

namespace tb {

constexpr int ROW_TILE = 16;  // feature rows per CTA in the row-tile kernels

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern std::atomic<long long> g_launches;  // diagnostics only (tb_launch_count)
inline void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
inline int launch_status() { return cudaGetLastError() == cudaSuccess ? TB_OK : TB_ERR_LAUNCH; }

int check_dims_host(const TbDims* d);

// tensor-core polyline encoder (tb_tc_kernels.cu)
constexpr int MAP_TC_MAX_CTA = 148;
size_t map_tc_scratch_bytes(int n_cta);
int launch_map_polyline_tc(const TbDims& d, const TbSceneIn& in, const float* packed, float* x0_scratch, int n_cta,
                           float* pl_feature, uint8_t* pl_valid, cudaStream_t st);

// the packed parameter buffer = [fp32 blob | pad to 1 KB | tensor-core blocks]
inline size_t tc_blob_offset_bytes() { return ((size_t)TB_PACKED_FLOATS * sizeof(float) + 1023) & ~(size_t)1023; }
inline const unsigned char* tc_blob(const float* packed) { return reinterpret_cast<const unsigned char*>(packed) + tc_blob_offset_bytes(); }

}  // namespace tb
