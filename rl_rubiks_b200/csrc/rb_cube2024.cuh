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
// Register LUT.  One move on a cubie-major state is 20 lookups new[j] = lut[a][kind(j)][s[j]] in a 24-entry byte
// table.  Done through shared memory that is 20 data-dependent LDS.U8 per state with random bank conflicts (the first
// version of these kernels: 25 % of the HBM roofline).  Here the two 24-byte rows of the state's action sit in 12
// registers and PRMT is the lookup.  For a word of 4 cubie values the selector nibbles are s & 15: bit 3 of a PRMT selector
// nibble means "replicate the sign of the selected byte", and every table entry is < 24, so a lookup in entries 0-7 / 16-23
// (selector s & 15) returns 0x00 exactly for the values 8-15, and a lookup in entries 8-15 with the selector's bit 3 flipped
// returns 0x00 for everything but the values 8-15 -- the sign mode does the range masking.  Only the choice between entries
// 0-7 and 16-23 (bit 4 of s) needs a byte mask: 10 ALU-pipe instructions per 4 lookups, no memory access.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt_r(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}

struct LutSel {              // per state word, independent of the action: reusable for all 12 actions in expand12
	uint32_t sel, selx, m2;
};
__device__ __forceinline__ LutSel lut_sel(uint32_t w) {
	LutSel q;
	const uint32_t y = (w & 0x0f0f0f0fu) | ((w >> 4) & 0xf0f0f0f0u);  // byte 0 = s0&15 | (s1&15)<<4, byte 2 = s2&15 | (s3&15)<<4
	q.sel = prmt_r(y, y, 0x3320u);                             // low 16 bits = the four selector nibbles (PRMT reads no more)
	q.selx = q.sel ^ 0x8888u;                                  // bit 3 of every nibble flipped: live for the values 8-15 only
	q.m2 = prmt_r(w * 8u, 0u, 0xba98u);                        // 0xff where bit 4 of s is set (entries 16-23)
	return q;
}
// row: the 24-byte table as 6 words (entries 4k..4k+3 in word k)
__device__ __forceinline__ uint32_t lut24(const uint32_t* __restrict__ row, const LutSel& q) {
	const uint32_t c0 = prmt_r(row[0], row[1], q.sel), c1 = prmt_r(row[2], row[3], q.selx), c2 = prmt_r(row[4], row[5], q.sel);
	const uint32_t r = c0 | c1;                                // s < 16: the entry (the other lookup gave 0); s >= 16: c0 is stale
	return (c2 & q.m2) | (r & ~q.m2);
}
// Row fetch.  Lanes of a warp hold different actions, and the 64-byte row pairs of the plain LUT all start in bank
// group 0 or 4: fetching them with LDS.128 is a 4-way bank conflict.  The kernels below therefore stage the LUT as
// s_rows[a][k][c] (k = 16-byte chunk of the row pair, c = copy 0..7): lane l reads copy l % 8, which lives in bank group
// l % 8 whatever its action is -- every fetch is conflict free (6 KB of shared memory).
constexpr int kRowsX8Vec = 12 * 4 * 8;          // uint4 elements
__device__ __forceinline__ void stage_rows_x8(uint4* s_rows) {
	for (int i = threadIdx.x; i < kRowsX8Vec; i += blockDim.x)
		s_rows[i] = reinterpret_cast<const uint4*>(g_lut2024)[i >> 3];        // i = (a*4 + k)*8 + c  ->  chunk a*4 + k
}
// The 12 row words (corner row, edge row) of action a.
__device__ __forceinline__ void load_rows(const uint4* s_rows, uint32_t a, int lane, uint32_t (&rc)[6], uint32_t (&re)[6]) {
	const uint4* p = s_rows + a * 32u + (lane & 7);
	const uint4 c0 = p[0], c1 = p[8], e0 = p[16], e1 = p[24];
	rc[0] = c0.x; rc[1] = c0.y; rc[2] = c0.z; rc[3] = c0.w; rc[4] = c1.x; rc[5] = c1.y;
	re[0] = e0.x; re[1] = e0.y; re[2] = e0.z; re[3] = e0.w; re[4] = e1.x; re[5] = e1.y;
}
__device__ __forceinline__ void move2024_reg(const uint32_t (&rc)[6], const uint32_t (&re)[6], uint32_t (&w)[5]) {
	w[0] = lut24(rc, lut_sel(w[0]));
	w[1] = lut24(rc, lut_sel(w[1]));
	w[2] = lut24(re, lut_sel(w[2]));
	w[3] = lut24(re, lut_sel(w[3]));
	w[4] = lut24(re, lut_sel(w[4]));
}

// ---------------------------------------------------------------------------------------------
// State streaming.  32 consecutive states are 640 contiguous bytes = 160 words: a warp moves them with five fully
// coalesced 128-byte accesses (lane l takes words l, l+32, ...), transposes through 640 bytes of its own shared memory
// (thread t owns words 5t..5t+4; stride 5 is coprime with the 32 banks: conflict free) and needs no block barrier.
// Requires 4-byte aligned state arrays; the tile kernels below (`*_any`) take everything else.
// ---------------------------------------------------------------------------------------------
constexpr int kWarpsPerBlock = kThreads / 32;

__device__ __forceinline__ void warp_load_states(uint32_t* tile, const uint32_t* __restrict__ src, int words, int lane) {
#pragma unroll
	for (int k = 0; k < 5; ++k) {
		const int i = lane + 32 * k;
		if (i < words) tile[i] = __ldcs(src + i);
	}
	__syncwarp();
}
__device__ __forceinline__ void warp_store_states(uint32_t* __restrict__ dst, const uint32_t* tile, int words, int lane) {
	__syncwarp();
#pragma unroll
	for (int k = 0; k < 5; ++k) {
		const int i = lane + 32 * k;
		if (i < words) __stcs(dst + i, tile[i]);
	}
	__syncwarp();
}

// multi_rotate: one move per state.  HBM-bound: 20 B in + 1-2 B action + 20 B out per state.  The loads of a warp's next
// chunk are issued before the current chunk is computed (register double buffer), so every warp always has 640 B in flight.
__global__ void __launch_bounds__(kThreads)
k_multi_rotate(const int8_t* __restrict__ in, const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs,
               int8_t* __restrict__ out, int64_t n) {
	__shared__ uint4 s_rows[kRowsX8Vec];
	__shared__ __align__(16) uint32_t s_tile[kWarpsPerBlock][160];
	stage_rows_x8(s_rows);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t* tile = s_tile[wib];
	const int64_t n_chunks = (n + 31) / 32, n_full = n / 32, stride = (int64_t)gridDim.x * kWarpsPerBlock;
	uint32_t nxt[5], nf = 0, nd = 0;
	// Only the last chunk can be ragged: whole chunks (warp-uniform test) run without the per-word bounds predicates, which were a
	// quarter of the ALU-pipe instructions of this ALU-bound kernel.
	auto fetch = [&](int64_t c) {                              // chunk c -> registers (words lane, lane+32, ...; action bytes)
		const int64_t base = c * 32;
		const uint32_t* src = reinterpret_cast<const uint32_t*>(in + base * 20);
		if (c < n_full) {
#pragma unroll
			for (int k = 0; k < 5; ++k) nxt[k] = __ldcs(src + lane + 32 * k);
			nf = faces[base + lane];
			nd = dirs ? dirs[base + lane] : 0u;
		} else {
			const int cnt = (int)(n - base);
#pragma unroll
			for (int k = 0; k < 5; ++k) nxt[k] = lane + 32 * k < cnt * 5 ? __ldcs(src + lane + 32 * k) : 0u;
			nf = lane < cnt ? faces[base + lane] : 0u;
			nd = (dirs && lane < cnt) ? dirs[base + lane] : 0u;
		}
	};
	int64_t c = (int64_t)blockIdx.x * kWarpsPerBlock + wib;
	if (c < n_chunks) fetch(c);
	for (; c < n_chunks; c += stride) {
		const int64_t base = c * 32;
		const bool full = c < n_full;
		const int cnt = full ? 32 : (int)(n - base);
		const uint32_t a = dirs ? rb_action_of(nf, nd) : rb_clamp_action(nf);
#pragma unroll
		for (int k = 0; k < 5; ++k) tile[lane + 32 * k] = nxt[k];
		if (c + stride < n_chunks) fetch(c + stride);
		__syncwarp();
		if (full || lane < cnt) {
			uint32_t w[5], rc[6], re[6];
#pragma unroll
			for (int k = 0; k < 5; ++k) w[k] = tile[lane * 5 + k];
			load_rows(s_rows, a, lane, rc, re);
			move2024_reg(rc, re, w);
#pragma unroll
			for (int k = 0; k < 5; ++k) tile[lane * 5 + k] = w[k];
		}
		uint32_t* dst = reinterpret_cast<uint32_t*>(out + base * 20);
		if (full) {
			__syncwarp();
#pragma unroll
			for (int k = 0; k < 5; ++k) __stcs(dst + lane + 32 * k, tile[lane + 32 * k]);
			__syncwarp();
		} else {
			warp_store_states(dst, tile, cnt * 5, lane);
		}
	}
}

// multi_is_solved: 20 B in + 1 B out per state.
__global__ void __launch_bounds__(kThreads)
k_is_solved(const int8_t* __restrict__ in, uint8_t* __restrict__ flags, int64_t n) {
	__shared__ __align__(16) uint32_t s_tile[kWarpsPerBlock][160];
	const uint32_t* sv = reinterpret_cast<const uint32_t*>(g_solved2024);
	const uint32_t s0 = sv[0], s1 = sv[1], s2 = sv[2], s3 = sv[3], s4 = sv[4];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t* tile = s_tile[wib];
	const int64_t n_chunks = (n + 31) / 32;
	for (int64_t c = (int64_t)blockIdx.x * kWarpsPerBlock + wib; c < n_chunks; c += (int64_t)gridDim.x * kWarpsPerBlock) {
		const int64_t base = c * 32;
		const int cnt = (int)min((int64_t)32, n - base);
		warp_load_states(tile, reinterpret_cast<const uint32_t*>(in + base * 20), cnt * 5, lane);
		if (lane < cnt) {
			const uint32_t* w = tile + lane * 5;
			flags[base + lane] = (w[0] == s0) & (w[1] == s1) & (w[2] == s2) & (w[3] == s3) & (w[4] == s4);
		}
		__syncwarp();
	}
}

// ---------------------------------------------------------------------------------------------
// Any-alignment versions (byte views at odd offsets): block tiles staged with whatever vector width the pointer allows.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_multi_rotate_any(const int8_t* __restrict__ in, const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs,
                   int8_t* __restrict__ out, int64_t n) {
	__shared__ uint4 s_rows[kRowsX8Vec];
	__shared__ __align__(16) uint32_t s_tile[kTile * 5];
	stage_rows_x8(s_rows);
	const int64_t n_tiles = (n + kTile - 1) / kTile;
	for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
		const int64_t base = tile * kTile;
		const int cnt = (int)min((int64_t)kTile, n - base);
		__syncthreads();                                      // previous tile fully stored / LUT staged
		rb_g2s(reinterpret_cast<uint8_t*>(s_tile), reinterpret_cast<const uint8_t*>(in) + base * 20, cnt * 20);
		__syncthreads();
		for (int i = threadIdx.x; i < cnt; i += kThreads) {
			const uint32_t a = dirs ? rb_action_of(faces[base + i], dirs[base + i]) : rb_clamp_action(faces[base + i]);
			uint32_t w[5], rc[6], re[6];
#pragma unroll
			for (int k = 0; k < 5; ++k) w[k] = s_tile[i * 5 + k];
			load_rows(s_rows, a, threadIdx.x & 31, rc, re);
			move2024_reg(rc, re, w);
#pragma unroll
			for (int k = 0; k < 5; ++k) s_tile[i * 5 + k] = w[k];
		}
		__syncthreads();
		rb_s2g(reinterpret_cast<uint8_t*>(out) + base * 20, reinterpret_cast<const uint8_t*>(s_tile), cnt * 20);
	}
}

__global__ void __launch_bounds__(kThreads)
k_is_solved_any(const int8_t* __restrict__ in, uint8_t* __restrict__ flags, int64_t n) {
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
__device__ __forceinline__ void warp_write_oh_row(float* __restrict__ row, uint32_t v, int lane, int pol) {
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
			rb_st_stream(dst + c, o, pol);
		}
	}
}

// bf16 rows (0x3F80 = 1.0, exact): 960 bytes = 60 x 16-byte chunks of 8 columns, lane l owns chunks l and l + 32.  Chunk c
// covers columns 8c..8c+7 inside cubie j = c/3 at offset 8*(c%3).  Same bits as the f32 row rounded to bf16.
typedef uint16_t rb_bf16;                       // raw bfloat16 bits at the C ABI
__device__ __forceinline__ void warp_write_oh_row(rb_bf16* __restrict__ row, uint32_t v, int lane, int pol) {
	uint4* dst = reinterpret_cast<uint4*>(row);
#pragma unroll
	for (int k = 0; k < 2; ++k) {
		const int c = lane + 32 * k;
		const int j = c / 3;
		const uint32_t vj = __shfl_sync(0xffffffffu, v, j < 20 ? j : 0);
		const uint32_t off = vj - 8u * (uint32_t)(c - 3 * j);
		if (c < kOhWidth / 8) {
			uint4 o;
			o.x = off == 0u ? 0x00003F80u : (off == 1u ? 0x3F800000u : 0u);
			o.y = off == 2u ? 0x00003F80u : (off == 3u ? 0x3F800000u : 0u);
			o.z = off == 4u ? 0x00003F80u : (off == 5u ? 0x3F800000u : 0u);
			o.w = off == 6u ? 0x00003F80u : (off == 7u ? 0x3F800000u : 0u);
			rb_st_stream(dst + c, o, pol);
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

// as_oh: HBM-write-bound, 20 B in + 1920 B out per state.  A block takes tiles of 256 consecutive states (5 words per
// thread, coalesced, fetched one tile ahead into registers so that the read latency -- several microseconds under a
// saturating write stream -- is paid once per 480 KB of output and hidden behind it); warp w then emits rows w, w+8, ...
// of the tile, so the block's eight warps sweep one contiguous region of the output.
template <typename OH>
__global__ void __launch_bounds__(kThreads)
k_as_oh(const int8_t* __restrict__ in, OH* __restrict__ oh, int64_t n, int pol) {
	__shared__ __align__(16) uint32_t s_tile[2][kThreads * 5];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t n_tiles = (n + kThreads - 1) / kThreads;
	uint32_t nxt[5];
	auto fetch = [&](int64_t t) {
		const int64_t base = t * kThreads;
		const int words = (int)min((int64_t)kThreads, n - base) * 5;
		const uint32_t* src = reinterpret_cast<const uint32_t*>(in + base * 20);
#pragma unroll
		for (int k = 0; k < 5; ++k) nxt[k] = (int)threadIdx.x + kThreads * k < words ? __ldcs(src + threadIdx.x + kThreads * k) : 0u;
	};
	int64_t t = blockIdx.x;
	if (t < n_tiles) fetch(t);
	for (int buf = 0; t < n_tiles; t += gridDim.x, buf ^= 1) {
		const int64_t base = t * kThreads;
		const int cnt = (int)min((int64_t)kThreads, n - base);
#pragma unroll
		for (int k = 0; k < 5; ++k) s_tile[buf][threadIdx.x + kThreads * k] = nxt[k];
		if (t + gridDim.x < n_tiles) fetch(t + gridDim.x);
		__syncthreads();                               // tile visible; the other buffer was last read two barriers ago
		const uint8_t* bytes = reinterpret_cast<const uint8_t*>(s_tile[buf]);
		for (int r = wib; r < cnt; r += kWarpsPerBlock) {
			const uint32_t v = lane < 20 ? (uint32_t)bytes[r * 20 + lane] : 0xffu;
			warp_write_oh_row(oh + (base + r) * kOhWidth, v, lane, pol);
		}
	}
}

// Any-alignment version: one warp per state, byte loads.
template <typename OH>
__global__ void __launch_bounds__(kThreads)
k_as_oh_any(const int8_t* __restrict__ in, OH* __restrict__ oh, int64_t n, int pol) {
	const int lane = threadIdx.x & 31;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t i = warp; i < n; i += n_warps) {
		const uint32_t v = warp_load_state(in + i * 20, lane);
		warp_write_oh_row(oh + i * kOhWidth, v, lane, pol);
	}
}

// ---------------------------------------------------------------------------------------------
// expand12 (+ fused one-hot + solved flags).  One warp per parent, 12 children each.
// Algorithmic bytes per parent: 20 in + 12*(20 + 1920 + 1) out.
// ---------------------------------------------------------------------------------------------
template <typename OH>
__device__ __forceinline__ void warp_expand12(const uint8_t* s_lut, uint32_t v, int lane, int64_t parent,
                                              int8_t* __restrict__ children, OH* __restrict__ children_oh,
                                              uint8_t* __restrict__ solved, int pol) {
#pragma unroll 4
	for (uint32_t a = 0; a < 12; ++a) {
		const uint32_t c = warp_move(s_lut, a, v, lane);
		const int64_t row = parent * 12 + a;
		if (children && lane < 20) children[row * 20 + lane] = (int8_t)c;
		if (children_oh) warp_write_oh_row(children_oh + row * kOhWidth, c, lane, pol);
		if (solved) {
			const bool s = warp_is_solved(c, lane);
			if (lane == 0) solved[row] = s;
		}
	}
}

template <typename OH>
__global__ void __launch_bounds__(kThreads)
k_expand12(const int8_t* __restrict__ in, int8_t* __restrict__ children, OH* __restrict__ children_oh,
           uint8_t* __restrict__ solved, int64_t n, int pol) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	rb_stage_lut2024(s_lut);
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t i = warp; i < n; i += n_warps) {
		const uint32_t v = warp_load_state(in + i * 20, lane);
		warp_expand12(s_lut, v, lane, i, children, children_oh, solved, pol);
	}
}

// ---------------------------------------------------------------------------------------------
// expand12, states (+ solved flags) only: 20 B in, 12 x 20 B out per parent, no one-hot to hide behind, so the
// warp-per-parent kernel above (20 active lanes, byte stores) reaches only 21 % of the roofline.  Here a thread owns a
// parent: the selector / mask part of the register LUT is computed once per state word and reused for all 12 actions,
// whose rows are constant-bank operands (every lane applies the same action).  The 240 output bytes of a parent are
// contiguous; a warp stages its 32 x 240 B in shared memory (row pitch 272 B: 16-byte aligned and conflict free for
// STS.128) and writes them out as fully coalesced 16-byte vectors.  Needs 16-byte aligned `children`, 4-byte `in`.
// ---------------------------------------------------------------------------------------------
constexpr int kExpThreads = 128;                    // 4 warps x 8.5 KB staging
constexpr int kExpPitch = 68;                       // words per parent row in shared memory (60 used)

__global__ void __launch_bounds__(kExpThreads)
k_expand12_states(const int8_t* __restrict__ in, int8_t* __restrict__ children, uint8_t* __restrict__ solved, int64_t n, int pol) {
	__shared__ __align__(16) uint32_t s_out[kExpThreads / 32][32 * kExpPitch];
	__shared__ __align__(16) uint32_t s_in[kExpThreads / 32][160];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t* tile = s_in[wib];
	uint32_t* stage = s_out[wib];
	const uint32_t* sv = reinterpret_cast<const uint32_t*>(g_solved2024);
	const uint32_t v0 = sv[0], v1 = sv[1], v2 = sv[2], v3 = sv[3], v4 = sv[4];
	const int64_t n_chunks = (n + 31) / 32;
	constexpr int kW = kExpThreads / 32;
	for (int64_t c = (int64_t)blockIdx.x * kW + wib; c < n_chunks; c += (int64_t)gridDim.x * kW) {
		const int64_t base = c * 32;
		const int cnt = (int)min((int64_t)32, n - base);
		warp_load_states(tile, reinterpret_cast<const uint32_t*>(in + base * 20), cnt * 5, lane);
		if (lane < cnt) {
			LutSel q[5];
#pragma unroll
			for (int k = 0; k < 5; ++k) q[k] = lut_sel(tile[lane * 5 + k]);
			uint32_t flags[3] = {0u, 0u, 0u};
			uint4* row = reinterpret_cast<uint4*>(stage + lane * kExpPitch);
#pragma unroll
			for (int g = 0; g < 3; ++g) {                      // 4 children = 20 words = 5 vectors at a time
				uint32_t w[20];
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const int a = 4 * g + j;
					const uint32_t* rc = c_lut2024 + a * 16;
					const uint32_t* re = rc + 8;
					w[5 * j + 0] = lut24(rc, q[0]);
					w[5 * j + 1] = lut24(rc, q[1]);
					w[5 * j + 2] = lut24(re, q[2]);
					w[5 * j + 3] = lut24(re, q[3]);
					w[5 * j + 4] = lut24(re, q[4]);
					const bool ok = (w[5 * j] == v0) & (w[5 * j + 1] == v1) & (w[5 * j + 2] == v2) & (w[5 * j + 3] == v3) & (w[5 * j + 4] == v4);
					flags[g] |= (ok ? 1u : 0u) << (8 * j);
				}
#pragma unroll
				for (int v = 0; v < 5; ++v) row[5 * g + v] = make_uint4(w[4 * v], w[4 * v + 1], w[4 * v + 2], w[4 * v + 3]);
			}
			if (solved) {                                      // 12 flag bytes per parent, 4-byte aligned (12 * index)
				uint32_t* f = reinterpret_cast<uint32_t*>(solved + (base + lane) * 12);
				f[0] = flags[0]; f[1] = flags[1]; f[2] = flags[2];
			}
		}
		__syncwarp();
		if (children) {
			uint4* dst = reinterpret_cast<uint4*>(children + base * 240);
			const int n_vec = cnt * 15;
			for (int i = lane; i < n_vec; i += 32) {
				const int p = (i * 2185) >> 15, r = i - 15 * p;    // i / 15 for i < 480
				rb_st_stream(dst + i, reinterpret_cast<const uint4*>(stage + p * kExpPitch)[r], pol);
			}
		}
		__syncwarp();
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
template <bool kChildren, typename OH>
__global__ void __launch_bounds__(kThreads)
k_sequence(const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs, int games, int depth,
           int with_solved, int chunk, int8_t* __restrict__ states, OH* __restrict__ oh,
           uint8_t* __restrict__ solved_states, int8_t* __restrict__ children, OH* __restrict__ children_oh,
           uint8_t* __restrict__ solved_children, int pol) {
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
			if (oh) warp_write_oh_row(oh + row * kOhWidth, v, lane, pol);
			if (solved_states) {
				const bool s = warp_is_solved(v, lane);
				if (lane == 0) solved_states[row] = s;
			}
			if (kChildren) warp_expand12(s_lut, v, lane, row, children, children_oh, solved_children, pol);
		}
	}
}

// ---------------------------------------------------------------------------------------------
// sequence_scramble without one-hot (states and/or solved flags only: 21 B per move).  With no 1920-byte row to hide
// behind, the warp-per-game kernel above (20 of 32 lanes, byte stores) reaches 4 % of the roofline; here a thread owns a
// game and a warp 32 consecutive games.  The state stays cubie-major in 5 registers, a move is the register LUT, the
// actions of the next 8 moves are prefetched while the current 8 are applied, and every 8 positions the warp flushes
// 32 x 160 contiguous bytes per game from shared memory (pitch 41 words: conflict free) in coalesced words.
// ---------------------------------------------------------------------------------------------
constexpr int kSeqGroup = 8;                    // positions staged per flush
constexpr int kSeqPitch = kSeqGroup * 5 + 1;    // words per game in the staging tile

__global__ void __launch_bounds__(kThreads, 3)
k_sequence_states(const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs, int games, int depth, int with_solved,
                  int8_t* __restrict__ states, uint8_t* __restrict__ solved_states) {
	__shared__ uint4 s_rows[kRowsX8Vec];
	__shared__ uint32_t s_stage[kWarpsPerBlock][32 * kSeqPitch];
	stage_rows_x8(s_rows);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t* stage = s_stage[wib];
	const uint32_t* sv = reinterpret_cast<const uint32_t*>(g_solved2024);
	const uint32_t v0 = sv[0], v1 = sv[1], v2 = sv[2], v3 = sv[3], v4 = sv[4];
	const int n_warps_games = (games + 31) / 32;
	const int moves = depth - with_solved;                     // action rows used: 0 .. moves-1
	const int64_t game_pitch = (int64_t)depth * 5;
	// Flush of a full group from a full warp of games: 1280 words, word i = lane + 32 j belongs to game i / 40, offset i % 40.
	// Five steps advance i by 160 = 4 games exactly, so (game, offset) of steps u = 0..4 are per-lane constants and step
	// j = 5 t + u is game 4 t + (that game): no division and no index arithmetic beyond one multiply-add per word in the loop (the
	// general form below cost ~10 instructions per word, a quarter of the kernel's instruction count).
	int64_t goff[5];
	int soff[5];
#pragma unroll
	for (int u = 0; u < 5; ++u) {
		const int i = lane + 32 * u, fg = i / (kSeqGroup * 5), fk = i - fg * (kSeqGroup * 5);
		goff[u] = fg * game_pitch + fk;
		soff[u] = fg * kSeqPitch + fk;
	}
	for (int wg = blockIdx.x * kWarpsPerBlock + wib; wg < n_warps_games; wg += gridDim.x * kWarpsPerBlock) {
		const int g0 = wg * 32, g = g0 + lane, cnt = min(32, games - g0);
		const bool live = lane < cnt;
		uint32_t w[5] = {v0, v1, v2, v3, v4};
		uint32_t act[kSeqGroup];
		const uint8_t* fcol = faces + g;                        // this lane's game: element m is fcol[m * games]
		const uint8_t* dcol = dirs ? dirs + g : nullptr;
		auto fetch = [&](int m0) {                             // actions of moves m0 .. m0+7 of this lane's game (coalesced over lanes)
			if (live && m0 >= 0 && m0 + kSeqGroup <= moves) {      // whole group inside the sequence: no per-move bounds
#pragma unroll
				for (int k = 0; k < kSeqGroup; ++k) {
					const int64_t off = (int64_t)(m0 + k) * games;
					act[k] = dcol ? rb_action_of(fcol[off], dcol[off]) : rb_clamp_action(fcol[off]);
				}
			} else {
#pragma unroll
				for (int k = 0; k < kSeqGroup; ++k) {
					const int m = m0 + k;
					uint32_t a = 12u;
					if (live && m >= 0 && m < moves) {
						const int64_t off = (int64_t)m * games;
						a = dcol ? rb_action_of(fcol[off], dcol[off]) : rb_clamp_action(fcol[off]);
					}
					act[k] = a;
				}
			}
		};
		// position d holds the state after d + 1 - with_solved moves: with_solved shifts the moves by one position
		fetch(-with_solved);
		for (int d0 = 0; d0 < depth; d0 += kSeqGroup) {
			uint32_t cur[kSeqGroup];
#pragma unroll
			for (int k = 0; k < kSeqGroup; ++k) cur[k] = act[k];
			if (d0 + kSeqGroup < depth) fetch(d0 + kSeqGroup - with_solved);
			uint32_t flags_lo = 0, flags_hi = 0;
			const int npos = min(kSeqGroup, depth - d0);
#pragma unroll
			for (int k = 0; k < kSeqGroup; ++k) {
				if (k < npos) {
					if (cur[k] < 12u) {                        // 12 = no move (the leading solved position / padding)
						uint32_t rc[6], re[6];
						load_rows(s_rows, cur[k], lane, rc, re);
						move2024_reg(rc, re, w);
					}
#pragma unroll
					for (int q = 0; q < 5; ++q) stage[lane * kSeqPitch + k * 5 + q] = w[q];
					if (solved_states) {
						const uint32_t ok = (w[0] == v0) & (w[1] == v1) & (w[2] == v2) & (w[3] == v3) & (w[4] == v4);
						if (k < 4) flags_lo |= ok << (8 * k); else flags_hi |= ok << (8 * (k - 4));
					}
				}
			}
			__syncwarp();
			if (states) {
				uint32_t* out_w = reinterpret_cast<uint32_t*>(states) + ((int64_t)g0 * depth + d0) * 5;
				if (npos == kSeqGroup && cnt == 32) {              // full group, full warp of games: the hoisted indices
#pragma unroll 2
					for (int t8 = 0; t8 < 8; ++t8) {
						uint32_t* o = out_w + (int64_t)(4 * t8) * game_pitch;
						const uint32_t* s = stage + 4 * t8 * kSeqPitch;
#pragma unroll
						for (int u = 0; u < 5; ++u) __stcs(o + goff[u], s[soff[u]]);
					}
				} else {
					const int per_game = npos * 5;
					for (int i = lane; i < cnt * per_game; i += 32) {
						const int gi = i / per_game, k = i - gi * per_game;
						out_w[gi * game_pitch + k] = stage[gi * kSeqPitch + k];
					}
				}
			}
			if (solved_states && live) {
				uint8_t* f = solved_states + (int64_t)g * depth + d0;
#pragma unroll
				for (int k = 0; k < kSeqGroup; ++k)
					if (k < npos) f[k] = (uint8_t)(((k < 4 ? flags_lo : flags_hi) >> (8 * (k & 3))) & 1u);
			}
			__syncwarp();
		}
	}
}

// ---------------------------------------------------------------------------------------------
// compose: states[i] <- "start[i], then the move sequence whose from-solved result is states[i]".  A sequence maps
// (position, orientation) pairs; its from-solved state lists the images of (p, 0) for every home position p, and
// orientation is additive in the twist / flip labelling of rb_scramble_macro.cuh, so the image of cubie value 3p+o is
// position p' = seq[p] / 3 with twist(p', seq[p] % 3) + twist(p, o).  This lets a scramble from arbitrary start
// states use the macro-move kernel.  Warp per state, lane = cubie, in place over `states`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t twist_of(uint32_t pos, uint32_t ori) {
	const bool neg = (0xa5u >> pos) & 1u;                 // positions 0, 2, 5, 7
	return neg ? (ori ? 3u - ori : 0u) : ori;
}
__global__ void __launch_bounds__(kThreads)
k_compose(int8_t* __restrict__ states, const int8_t* __restrict__ start, int64_t n) {
	const int lane = threadIdx.x & 31;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t i = warp; i < n; i += n_warps) {
		const uint32_t seq = warp_load_state(states + i * 20, lane);      // lane p < 8: image of corner position p; lane 8+p: of edge position p
		const uint32_t s = warp_load_state(start + i * 20, lane) & 31u;
		const bool corner = lane < 8;
		const uint32_t p = corner ? (s * 11u) >> 5 : s >> 1;               // s / 3 for s < 32
		const uint32_t o = corner ? s - 3u * p : s & 1u;
		const uint32_t img = __shfl_sync(0xffffffffu, seq, corner ? min(p, 7u) : 8u + min(p, 11u)) & 31u;
		uint32_t out;
		if (corner) {
			const uint32_t p2 = (img * 11u) >> 5, o2 = img - 3u * p2;
			uint32_t t = twist_of(p2 & 7u, o2) + twist_of(p & 7u, o);
			t = t >= 3u ? t - 3u : t;
			out = 3u * p2 + twist_of(p2 & 7u, t);                          // twist <-> orientation is an involution per position
		} else {
			out = img ^ o;
		}
		if (lane < 20) states[i * 20 + lane] = (int8_t)out;
	}
}

}  // namespace rb2024
