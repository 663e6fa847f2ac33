// ADI target assembly and loss weights (reference: librubiks/train.py:292-296, 313-333).
#pragma once
#include "rb_common.cuh"

namespace rbadi {

constexpr int kThreads = 256;

// One thread per scrambled state: 12 child values (48 B, three 16-byte loads), 12 child solved flags.
// values + reward is a single f32 add per child (IEEE, same as torch's `values += rewards`); argmax keeps the
// FIRST maximum; a NaN is treated as the maximum (torch.argmax semantics).
__global__ void __launch_bounds__(kThreads)
k_targets(const float* __restrict__ values, const uint8_t* __restrict__ solved_children,
          const uint8_t* __restrict__ solved_states, int64_t n, int depth, int method,
          int64_t* __restrict__ policy, float* __restrict__ value, float* __restrict__ weights, double alpha, double ws) {
	const float win = method == RB_REWARD_REWARD0 ? 0.f : 1.f;
	const double us = (double)n;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		float v[12];
		const float4* vp = reinterpret_cast<const float4*>(values + i * 12);   // 48 B rows: 16-byte aligned
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const float4 q = vp[k];
			v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
		}
		const uint32_t* sp = reinterpret_cast<const uint32_t*>(solved_children + i * 12);  // 12 B rows: 4-byte aligned
		const uint32_t f0 = sp[0], f1 = sp[1], f2 = sp[2];
		int best = 0;
		float bv = 0.f;
#pragma unroll
		for (int a = 0; a < 12; ++a) {
			const uint32_t fw = a < 4 ? f0 : (a < 8 ? f1 : f2);
			const bool s = (fw >> (8 * (a & 3))) & 0xffu;
			const float x = v[a] + (s ? win : -1.f);
			if (a == 0) { bv = x; }
			else if (!(bv != bv) && (x > bv || x != x)) { bv = x; best = a; }
		}
		if (method == RB_REWARD_LAPANFIX) { if (solved_states[i]) bv = 0.f; }
		else if (method == RB_REWARD_SCHULTZFIX) { if (i % depth == 0) bv = 0.f; }
		policy[i] = best;
		value[i] = bv;
		if (weights) {                                 // loss weights fused in (same arithmetic as k_loss_weights below)
			const double w = 1.0 / (double)(1 + (int)(i % depth));
			const double x = __dadd_rn(__ddiv_rn(__dmul_rn(1.0 - alpha, w), ws), __ddiv_rn(__dmul_rn(alpha, 1.0), us));
			weights[i] = (float)__dmul_rn(x, __dadd_rn(ws, us));
		}
	}
}

// f64 throughout, one rounding to f32 at the end, operation order exactly as train.py:333:
// ((1-alpha) * w / ws + alpha * 1 / us) * (ws + us)
__global__ void __launch_bounds__(kThreads)
k_loss_weights(float* __restrict__ out, int64_t n, int depth, double alpha, double ws) {
	const double us = (double)n;
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
		const double w = 1.0 / (double)(1 + (int)(i % depth));
		const double x = __dadd_rn(__ddiv_rn(__dmul_rn(1.0 - alpha, w), ws), __ddiv_rn(__dmul_rn(alpha, 1.0), us));
		out[i] = (float)__dmul_rn(x, __dadd_rn(ws, us));
	}
}

// numpy's pairwise float64 summation (numpy/_core/src/umath/loops_utils.h.src, `DOUBLE_pairwise_sum`), which is
// what `weighted.sum()` at train.py:332 runs on the contiguous tiled array.
static double pairwise_sum(const double* a, int64_t n) {
	if (n < 8) {
		double res = 0.;
		for (int64_t i = 0; i < n; ++i) res += a[i];
		return res;
	}
	if (n <= 128) {
		double r[8];
		for (int k = 0; k < 8; ++k) r[k] = a[k];
		int64_t i;
		for (i = 8; i < n - (n % 8); i += 8)
			for (int k = 0; k < 8; ++k) r[k] += a[i + k];
		double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
		for (; i < n; ++i) res += a[i];
		return res;
	}
	int64_t n2 = n / 2;
	n2 -= n2 % 8;
	return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
}

}  // namespace rbadi
