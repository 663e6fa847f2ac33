// 20x24 representation kernels: multi_rotate, multi_is_solved, as_oh, expand12, scramble,
// sequence_scramble and the fused ADI generator.
//
// Data layout in HBM (the reference's own): state = int8[20] (8 corners, value 3*pos+ori, then 12 edges,
// value 2*pos+ori), n states contiguous (20 B pitch); one-hot = f32[480] per state (1920 B pitch).
// A state is 5 x 32-bit words: words 0-1 corners, 2-4 edges.  20 B records are not a power of two, so the
// state-streaming kernels move tiles of TILE states (TILE*20 B, a multiple of 16) through shared memory with
// 16-byte coalesced accesses; in shared memory thread t reads words 5t..5t+4, and since gcd(5,32)=1 that is
// bank-conflict free.
#pragma once
#include "rb_common.cuh"

namespace rb2024 {

constexpr int kThreads = 256;
constexpr int kTile = 1024;             // states per tile: 20 KB of shared memory
constexpr int kOhWidth = 480;
constexpr int kOhVec = kOhWidth / 4;    // 120 float4 per one-hot row

// ---------------------------------------------------------------------------------------------
// multi_rotate: one move per state.  HBM-bound: 20 B in + 1-2 B action + 20 B out per state.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_multi_rotate(const int8_t* __restrict__ in, const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs,
               int8_t* __restrict__ out, int64_t n) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	__shared__ __align__(16) uint32_t s_tile[kTile * 5];
	rb_stage_lut2024(s_lut);
	const int64_t n_tiles = (n + kTile - 1) / kTile;
	for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const int64_t base = tile * kTile;
		const int cnt = (int)min((int64_t)kTile, n - base);
		__syncthreads();                                      // previous tile fully stored / LUT staged
		rb_g2s(reinterpret_cast<uint8_t*>(s_tile), reinterpret_cast<const uint8_t*>(in) + base * 20, cnt * 20);
		__syncthreads();
		for (int i = threadIdx.x; i < cnt; i += kThreads) {
			uint32_t a = dirs ? rb_action_of(faces[base + i], dirs[base + i]) : rb_clamp_action(faces[base + i]);
			uint32_t w[5];
#pragma unroll
			for (int k = 0; k < 5; ++k) w[k] = s_tile[i * 5 + k];
			rb_move2024(s_lut, a, w);
#pragma unroll
			for (int k = 0; k < 5; ++k) s_tile[i * 5 + k] = w[k];
		}
		__syncthreads();
		rb_s2g(reinterpret_cast<uint8_t*>(out) + base * 20, reinterpret_cast<const uint8_t*>(s_tile), cnt * 20);
	}
}

// ---------------------------------------------------------------------------------------------
// multi_is_solved
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_is_solved(const int8_t* __restrict__ in, uint8_t* __restrict__ flags, int64_t n) {
	__shared__ __align__(16) uint32_t s_tile[kTile * 5];
	const uint32_t* sv = reinterpret_cast<const uint32_t*>(g_solved2024);
	const uint32_t s0 = sv[0], s1 = sv[1], s2 = sv[2], s3 = sv[3], s4 = sv[4];
	const int64_t n_tiles = (n + kTile - 1) / kTile;
	for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const int64_t base = tile * kTile;
		const int cnt = (int)min((int64_t)kTile, n - base);
		__syncthreads();
		rb_g2s(reinterpret_cast<uint8_t*>(s_tile), reinterpret_cast<const uint8_t*>(in) + base * 20, cnt * 20);
		__syncthreads();
		for (int i = threadIdx.x; i < cnt; i += kThreads) {
			const uint32_t* w = s_tile + i * 5;
			flags[base + i] = (w[0] == s0) & (w[1] == s1) & (w[2] == s2) & (w[3] == s3) & (w[4] == s4);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative one-hot row.  The 1920-byte row is 120 x 16-byte chunks; lane l owns chunks l, l+32,
// l+64, l+96.  Chunk c covers columns 4c..4c+3, all inside cubie j = c/6 at offset 4*(c%6), so its four
// floats are (v_j - 4*(c%6) == 0,1,2,3).  `v` is the cubie value held by lane j (lanes 0..19); every
// element of the row is written (no zero-fill pass, no index tensors), one coalesced 512-byte store per step.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_write_oh_row(float* __restrict__ row, uint32_t v, int lane) {
	float4* dst = reinterpret_cast<float4*>(row);
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		const int c = lane + 32 * k;                 // chunk index
		const int j = c / 6;                         // owning cubie
		const uint32_t vj = __shfl_sync(0xffffffffu, v, j < 20 ? j : 0);
		const uint32_t off = vj - 4u * (uint32_t)(c - 6 * j);
		if (c < kOhVec) {
			float4 o;
			o.x = off == 0u ? 1.f : 0.f;
			o.y = off == 1u ? 1.f : 0.f;
			o.z = off == 2u ? 1.f : 0.f;
			o.w = off == 3u ? 1.f : 0.f;
			rb_st_stream(dst + c, o);
		}
	}
}

// Lane j < 20 holds cubie j of the state at `p` (int8[20]); other lanes hold 0xff.
__device__ __forceinline__ uint32_t warp_load_state(const int8_t* __restrict__ p, int lane) {
	return lane < 20 ? (uint32_t)(uint8_t)p[lane] : 0xffu;
}

__device__ __forceinline__ bool warp_is_solved(uint32_t v, int lane) {
	const bool ok = lane >= 20 || v == (uint32_t)g_solved2024[lane];
	return __all_sync(0xffffffffu, ok);
}

// One move, warp-cooperative: lane j looks up cubie j.  All 20 active lanes read the same two 24-byte rows,
// 6 words each in distinct banks: conflict free.
__device__ __forceinline__ uint32_t warp_move(const uint8_t* s_lut, uint32_t a, uint32_t v, int lane) {
	return lane < 20 ? (uint32_t)s_lut[a * 64u + (lane >= 8 ? 32u : 0u) + (v & 31u)] : 0xffu;
}

// as_oh: HBM-bound, 20 B in + 1920 B out per state.  One warp per state, grid-stride.
__global__ void __launch_bounds__(kThreads)
k_as_oh(const int8_t* __restrict__ in, float* __restrict__ oh, int64_t n) {
	const int lane = threadIdx.x & 31;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t i = warp; i < n; i += n_warps) {
		const uint32_t v = warp_load_state(in + i * 20, lane);
		warp_write_oh_row(oh + i * kOhWidth, v, lane);
	}
}

// ---------------------------------------------------------------------------------------------
// expand12 (+ fused one-hot + solved flags).  One warp per parent, 12 children each.
// Algorithmic bytes per parent: 20 in + 12*(20 + 1920 + 1) out.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_expand12(const uint8_t* s_lut, uint32_t v, int lane, int64_t parent,
                                              int8_t* __restrict__ children, float* __restrict__ children_oh,
                                              uint8_t* __restrict__ solved) {
#pragma unroll 4
	for (uint32_t a = 0; a < 12; ++a) {
		const uint32_t c = warp_move(s_lut, a, v, lane);
		const int64_t row = parent * 12 + a;
		if (children && lane < 20) children[row * 20 + lane] = (int8_t)c;
		if (children_oh) warp_write_oh_row(children_oh + row * kOhWidth, c, lane);
		if (solved) {
			const bool s = warp_is_solved(c, lane);
			if (lane == 0) solved[row] = s;
		}
	}
}

__global__ void __launch_bounds__(kThreads)
k_expand12(const int8_t* __restrict__ in, int8_t* __restrict__ children, float* __restrict__ children_oh,
           uint8_t* __restrict__ solved, int64_t n) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	rb_stage_lut2024(s_lut);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t i = warp; i < n; i += n_warps) {
		const uint32_t v = warp_load_state(in + i * 20, lane);
		warp_expand12(s_lut, v, lane, i, children, children_oh, solved);
	}
}

// ---------------------------------------------------------------------------------------------
// scramble: `depth` moves per cube, final state only (thread per cube, state in 5 registers).
// v1: byte-LUT lookups in shared memory; bound by shared-memory lookups, not HBM.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_scramble(const uint8_t* __restrict__ actions, int64_t stride_cube, int64_t stride_move,
           const int8_t* __restrict__ start, int8_t* __restrict__ out, int64_t n, int depth) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	__shared__ __align__(16) uint32_t s_tile[kThreads * 5];
	rb_stage_lut2024(s_lut);
	const int64_t n_tiles = (n + kThreads - 1) / kThreads;
	for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const int64_t base = tile * kThreads;
		const int cnt = (int)min((int64_t)kThreads, n - base);
		__syncthreads();
		if (start) rb_g2s(reinterpret_cast<uint8_t*>(s_tile), reinterpret_cast<const uint8_t*>(start) + base * 20, cnt * 20);
		__syncthreads();
		const int i = threadIdx.x;
		if (i < cnt) {
			uint32_t w[5];
			const uint32_t* src = start ? s_tile + i * 5 : reinterpret_cast<const uint32_t*>(g_solved2024);
#pragma unroll
			for (int k = 0; k < 5; ++k) w[k] = src[k];
			const uint8_t* ap = actions + (base + i) * stride_cube;
			if (stride_move == 1 && (stride_cube & 3) == 0 && (reinterpret_cast<uintptr_t>(actions) & 3u) == 0) {
				int m = 0;
				for (; m + 4 <= depth; m += 4) {
					const uint32_t a4 = *reinterpret_cast<const uint32_t*>(ap + m);
#pragma unroll
					for (int k = 0; k < 4; ++k) rb_move2024(s_lut, rb_clamp_action((a4 >> (8 * k)) & 0xffu), w);
				}
				for (; m < depth; ++m) rb_move2024(s_lut, rb_clamp_action(ap[m]), w);
			} else {
				for (int m = 0; m < depth; ++m) rb_move2024(s_lut, rb_clamp_action(ap[(int64_t)m * stride_move]), w);
			}
#pragma unroll
			for (int k = 0; k < 5; ++k) s_tile[i * 5 + k] = w[k];
		}
		__syncthreads();
		rb_s2g(reinterpret_cast<uint8_t*>(out) + base * 20, reinterpret_cast<const uint8_t*>(s_tile), cnt * 20);
	}
}

// ---------------------------------------------------------------------------------------------
// sequence_scramble and the fused ADI generator.  Work unit = (game, chunk of `chunk` consecutive depth
// positions); one warp per unit.  The warp replays the game's first moves to reach the chunk (cheap: one
// conflict-free shared-memory lookup per lane per move), then for every position in the chunk emits the
// state, its one-hot row, its solved flag and (ADI) the 12 children with their one-hot rows and flags.
// Row index of (game g, position d) is g*depth + d: game-major, depth-minor (cube.py:232).
// ---------------------------------------------------------------------------------------------
template <bool kChildren>
__global__ void __launch_bounds__(kThreads)
k_sequence(const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs, int games, int depth,
           int with_solved, int chunk, int8_t* __restrict__ states, float* __restrict__ oh,
           uint8_t* __restrict__ solved_states, int8_t* __restrict__ children, float* __restrict__ children_oh,
           uint8_t* __restrict__ solved_children) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	rb_stage_lut2024(s_lut);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int chunks_per_game = (depth + chunk - 1) / chunk;
	const int64_t n_units = (int64_t)games * chunks_per_game;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t u = warp; u < n_units; u += n_warps) {
		const int g = (int)(u / chunks_per_game);
		const int d0 = (int)(u % chunks_per_game) * chunk;
		const int d1 = min(depth, d0 + chunk);
		// Position d holds the state after `d + 1 - with_solved` moves (cube.py:228-231).
		uint32_t v = lane < 20 ? (uint32_t)g_solved2024[lane] : 0xffu;
		const int total = d1 - with_solved;               // moves needed for the last position of the chunk
		int applied = 0, buf0 = 0;
		uint32_t a_l = 0;                                 // lane l holds the action of move buf0 + l
		auto fetch = [&]() {
			const int m = buf0 + lane;
			a_l = 0;
			if (m < total) {
				const int64_t idx = (int64_t)m * games + g;
				a_l = dirs ? rb_action_of(faces[idx], dirs[idx]) : rb_clamp_action(faces[idx]);
			}
		};
		fetch();
		for (int d = d0; d < d1; ++d) {
			const int want = d + 1 - with_solved;
			while (applied < want) {
				if (applied - buf0 == 32) { buf0 = applied; fetch(); }
				v = warp_move(s_lut, __shfl_sync(0xffffffffu, a_l, applied - buf0), v, lane);
				++applied;
			}
			const int64_t row = (int64_t)g * depth + d;
			if (states && lane < 20) states[row * 20 + lane] = (int8_t)v;
			if (oh) warp_write_oh_row(oh + row * kOhWidth, v, lane);
			if (solved_states) {
				const bool s = warp_is_solved(v, lane);
				if (lane == 0) solved_states[row] = s;
			}
			if (kChildren) warp_expand12(s_lut, v, lane, row, children, children_oh, solved_children);
		}
	}
}

}  // namespace rb2024
