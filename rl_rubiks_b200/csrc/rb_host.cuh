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

// Packed actions -> action bytes.  Packed byte p = a0 + 13 * a1 holds moves (2k, 2k + 1) of a cube (a1 = 12: no second move, only
// in the last byte of an odd-depth row) -- the row index of the 2-move table (rb_get_macro_table).  Rows: packed [n][(depth + 1) / 2],
// actions [n][depth].  Even depth is one flat stream: 8 packed bytes in, 16 action bytes out per thread.
__global__ void __launch_bounds__(256)
k_unpack_actions(const uint8_t* __restrict__ packed, uint8_t* __restrict__ actions, int64_t n, int depth, int vec_ok) {
	const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const int pb = (depth + 1) >> 1;
	auto split = [](uint32_t w, uint32_t& lo, uint32_t& hi) {                 // 4 packed bytes -> 8 action bytes, two bytes per 16-bit lane
		const uint32_t ev = w & 0x00ff00ffu, od = (w >> 8) & 0x00ff00ffu;      // packed bytes (0, 2) and (1, 3)
		const uint32_t qe = ((ev * 79u) >> 10) & 0x003f003fu, qo = ((od * 79u) >> 10) & 0x003f003fu;   // p / 13 per lane (p * 79 < 2^16: no carry across lanes)
		const uint32_t pe = ev - 13u * qe + (qe << 8), po = od - 13u * qo + (qo << 8);                 // lanes (a, q): the two moves of a packed byte
		asm("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(lo) : "r"(pe), "r"(po));    // moves of packed bytes 0, 1
		asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(hi) : "r"(pe), "r"(po));    // moves of packed bytes 2, 3
	};
	if (vec_ok) {                                                            // even depth, 16-byte aligned pointers
		// 8 packed bytes in, 16 action bytes out per thread: both the loads (256 B per warp) and the stores (512 B) are contiguous
		const int64_t total = n * pb, nv = total >> 3;
		for (int64_t i = t0; i < nv; i += stride) {
			const uint2 v = __ldcs(reinterpret_cast<const uint2*>(packed) + i);
			uint4 o;
			split(v.x, o.x, o.y); split(v.y, o.z, o.w);
			__stcs(reinterpret_cast<uint4*>(actions) + i, o);
		}
		for (int64_t i = (nv << 3) + t0; i < total; i += stride) {
			const uint32_t p = packed[i], q = (p * 79u) >> 10;
			actions[2 * i] = (uint8_t)(p - 13u * q); actions[2 * i + 1] = (uint8_t)q;
		}
		return;
	}
	const int64_t total = n * pb;
	for (int64_t i = t0; i < total; i += stride) {
		const int64_t c = i / pb;
		const int k = (int)(i - c * pb);
		const uint32_t p = packed[i], q = (p * 79u) >> 10;
		actions[c * depth + 2 * k] = (uint8_t)(p - 13u * q);
		if (2 * k + 1 < depth) actions[c * depth + 2 * k + 1] = (uint8_t)q;
	}
}

// Device staging for the host-buffer entry points: two slots so that the H2D copy of chunk k+1, the kernel on
// chunk k and the D2H copy of chunk k-1 overlap.  Grown on demand, released by rbh_release().
struct Staging {
	void* buf[2][4] = {{nullptr, nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr, nullptr}};   // per slot: in, in2, out, scratch
	size_t bytes[4] = {0, 0, 0, 0};
	cudaStream_t stream[2] = {nullptr, nullptr};
	int device = -1;
};
static Staging g_stage;
static std::mutex g_stage_mu;

// Frees every buffer and stream on the device that owns them; pointers are nulled and sizes zeroed BEFORE anything can fail.
static void stage_drop(Staging& s) {
	int cur = -1;
	if (s.device >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != s.device) cudaSetDevice(s.device);
	for (int k = 0; k < 2; ++k) {
		for (int b = 0; b < 4; ++b) {
			void* p = s.buf[k][b];
			s.buf[k][b] = nullptr;
			if (p) cudaFree(p);
		}
		cudaStream_t st = s.stream[k];
		s.stream[k] = nullptr;
		if (st) cudaStreamDestroy(st);
	}
	for (int b = 0; b < 4; ++b) s.bytes[b] = 0;
	if (cur >= 0 && cur != s.device) cudaSetDevice(cur);
	s.device = -1;
}

static int stage_reserve(size_t in_bytes, size_t in2_bytes, size_t out_bytes, size_t scratch_bytes = 0) {
	int dev = 0;
	RB_CUDA(cudaGetDevice(&dev));
	Staging& s = g_stage;
	if (s.device != dev) {
		stage_drop(s);
		s.device = dev;
	}
	const size_t want[4] = {in_bytes, in2_bytes, out_bytes, scratch_bytes};
	for (int k = 0; k < 2; ++k)
		if (!s.stream[k]) RB_CUDA(cudaStreamCreateWithFlags(&s.stream[k], cudaStreamNonBlocking));
	for (int b = 0; b < 4; ++b) {
		if (want[b] <= s.bytes[b]) continue;
		s.bytes[b] = 0;                                    // a failed allocation leaves "nothing reserved", never a dangling pointer
		for (int k = 0; k < 2; ++k) {
			void* p = s.buf[k][b];
			s.buf[k][b] = nullptr;
			if (p) RB_CUDA(cudaFree(p));
		}
		for (int k = 0; k < 2; ++k) RB_CUDA(cudaMalloc(&s.buf[k][b], want[b]));
		s.bytes[b] = want[b];
	}
	return RB_OK;
}

// Waits for both staging streams (used on every exit path of an rbh_* call: no copy to or from the caller's host buffers may
// still be in flight when the call returns) and reports the first error.
static int stage_sync(int rc) {
	Staging& s = g_stage;
	for (int k = 0; k < 2; ++k) {
		if (!s.stream[k]) continue;
		const cudaError_t e = cudaStreamSynchronize(s.stream[k]);
		if (e != cudaSuccess && rc == RB_OK) rc = rb_fail(RB_ERR_CUDA, "%s: %s", "cudaStreamSynchronize(staging stream)", cudaGetErrorString(e));
	}
	return rc;
}

static int stage_release() {
	stage_drop(g_stage);
	return RB_OK;
}

}  // namespace rbh
