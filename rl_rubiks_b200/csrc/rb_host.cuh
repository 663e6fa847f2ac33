// Range check and the host-buffer (end-to-end) entry points.
#pragma once
#include "rb_common.cuh"
#include "rb_tables.cuh"
#include "rb_cube2024.cuh"
#include "rb_cube686.cuh"

namespace rbh {

__global__ void __launch_bounds__(256)
k_check_range(int rep, const uint8_t* __restrict__ states, int64_t n_state_bytes, const uint8_t* __restrict__ faces,
              const uint8_t* __restrict__ dirs, int64_t n_actions, int32_t* __restrict__ flag) {
	bool bad = false;
	const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	for (int64_t i = t0; i < n_actions; i += stride)
		bad |= dirs ? (faces[i] >= 6 || dirs[i] >= 2) : (faces[i] >= 12);
	if (rep == RB_REP_2024 && states)
		for (int64_t i = t0; i < n_state_bytes; i += stride) bad |= states[i] >= 24;
	if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(flag, 1);
}

// Device staging for the host-buffer entry points: two slots so that the H2D copy of chunk k+1, the kernel on
// chunk k and the D2H copy of chunk k-1 overlap.  Grown on demand, released by rbh_release().
struct Staging {
	void* in[2] = {nullptr, nullptr};
	void* in2[2] = {nullptr, nullptr};
	void* out[2] = {nullptr, nullptr};
	size_t in_bytes = 0, in2_bytes = 0, out_bytes = 0;
	cudaStream_t stream[2] = {nullptr, nullptr};
	int device = -1;
};
static Staging g_stage;
static std::mutex g_stage_mu;

static int stage_reserve(size_t in_bytes, size_t in2_bytes, size_t out_bytes) {
	int dev = 0;
	RB_CUDA(cudaGetDevice(&dev));
	Staging& s = g_stage;
	if (s.device != dev) {
		for (int k = 0; k < 2; ++k) {
			if (s.in[k]) cudaFree(s.in[k]);
			if (s.in2[k]) cudaFree(s.in2[k]);
			if (s.out[k]) cudaFree(s.out[k]);
			if (s.stream[k]) cudaStreamDestroy(s.stream[k]);
			s.in[k] = s.in2[k] = s.out[k] = nullptr;
			s.stream[k] = nullptr;
		}
		s.in_bytes = s.in2_bytes = s.out_bytes = 0;
		s.device = dev;
	}
	for (int k = 0; k < 2; ++k) {
		if (!s.stream[k]) RB_CUDA(cudaStreamCreateWithFlags(&s.stream[k], cudaStreamNonBlocking));
		if (in_bytes > s.in_bytes) { if (s.in[k]) cudaFree(s.in[k]); RB_CUDA(cudaMalloc(&s.in[k], in_bytes)); }
		if (in2_bytes > s.in2_bytes) { if (s.in2[k]) cudaFree(s.in2[k]); RB_CUDA(cudaMalloc(&s.in2[k], in2_bytes)); }
		if (out_bytes > s.out_bytes) { if (s.out[k]) cudaFree(s.out[k]); RB_CUDA(cudaMalloc(&s.out[k], out_bytes)); }
	}
	if (in_bytes > s.in_bytes) s.in_bytes = in_bytes;
	if (in2_bytes > s.in2_bytes) s.in2_bytes = in2_bytes;
	if (out_bytes > s.out_bytes) s.out_bytes = out_bytes;
	return RB_OK;
}

static int stage_release() {
	Staging& s = g_stage;
	for (int k = 0; k < 2; ++k) {
		if (s.in[k]) cudaFree(s.in[k]);
		if (s.in2[k]) cudaFree(s.in2[k]);
		if (s.out[k]) cudaFree(s.out[k]);
		if (s.stream[k]) cudaStreamDestroy(s.stream[k]);
		s.in[k] = s.in2[k] = s.out[k] = nullptr;
		s.stream[k] = nullptr;
	}
	s.in_bytes = s.in2_bytes = s.out_bytes = 0;
	s.device = -1;
	return RB_OK;
}

}  // namespace rbh
