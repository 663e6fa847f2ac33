// 6x8x6 representation kernels.  A state is int8[6][8][6] = 48 sticker slots x 6-byte colour one-hot
// (288 B, 72 words, 18 x 16 B).  Every move is a fixed permutation of the 48 slots that moves exactly 20 of
// them (reference: librubiks/cube/cube.py:330-361); the kernels move whole 6-byte records, so any int8
// content is carried bit-exactly, as the reference's fancy-index assignment does.
#pragma once
#include "rb_common.cuh"

namespace rb686 {

constexpr int kThreads = 256;
constexpr int kStateBytes = 288;
constexpr int kTile = 64;               // states per tile: 18 KB in + 18 KB out of shared memory
constexpr int kSlots = 48;

__device__ __forceinline__ void stage_perm(uint8_t* s_perm) {
	for (int i = threadIdx.x; i < 12 * 48 / 16; i += blockDim.x)
		reinterpret_cast<uint4*>(s_perm)[i] = reinterpret_cast<const uint4*>(g_perm686)[i];
}

// Copy one 6-byte sticker record inside shared memory (2-byte aligned).
__device__ __forceinline__ void copy_record(uint8_t* dst, const uint8_t* src) {
	const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
	uint16_t* d = reinterpret_cast<uint16_t*>(dst);
	const uint16_t a = s[0], b = s[1], c = s[2];
	d[0] = a; d[1] = b; d[2] = c;
}

// Halfword gather table.  A 6-byte sticker record is 3 halfwords and a state 144 of them; output word j of a move is the
// halfword pair (2j, 2j+1), each fetched from source halfword perm[a][h / 3] * 3 + h % 3.  s_tab[a][j] holds the two source
// halfword indices (< 144) as bytes.  28 of the 48 records do not move, so most lanes read consecutive halfwords.
__device__ __forceinline__ void stage_gather(uint16_t* s_tab) {
	for (int i = threadIdx.x; i < 12 * 72; i += blockDim.x) {
		const int a = i / 72, j = i - 72 * a, h0 = 2 * j, h1 = h0 + 1;
		const uint32_t s0 = g_perm686[a * kSlots + h0 / 3] * 3u + h0 % 3, s1 = g_perm686[a * kSlots + h1 / 3] * 3u + h1 % 3;
		s_tab[i] = (uint16_t)(s0 | (s1 << 8));
	}
}
// Output word `j` of action `a` applied to the state at `src` (144 halfwords in shared memory).
__device__ __forceinline__ uint32_t gather_word(const uint16_t* s_tab, const uint16_t* src, uint32_t a, uint32_t j) {
	const uint32_t pair = s_tab[a * 72u + j];
	return (uint32_t)src[pair & 0xffu] | ((uint32_t)src[pair >> 8] << 16);
}

// multi_rotate: 288 B in + action + 288 B out per state.  A warp takes 4 states (288 words = 9 full coalesced rounds)
// into its own 1152 bytes of shared memory and writes every output word straight from the gather to global memory.
constexpr int kMrStates = 4;
__global__ void __launch_bounds__(kThreads)
k_multi_rotate(const int8_t* __restrict__ in, const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs,
               int8_t* __restrict__ out, int64_t n) {
	__shared__ uint16_t s_tab[12 * 72];
	__shared__ __align__(16) uint32_t s_buf[kThreads / 32][kMrStates * 72];
	stage_gather(s_tab);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t* buf = s_buf[wib];
	const int64_t n_groups = (n + kMrStates - 1) / kMrStates, stride = (int64_t)gridDim.x * (kThreads / 32);
	uint32_t nxt[9], a_n = 0;
	auto fetch = [&](int64_t g) {                              // group g -> registers, issued one iteration ahead
		const int64_t base = g * kMrStates;
		const int cnt = (int)min((int64_t)kMrStates, n - base), words = cnt * 72;
		const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + base * 72;
#pragma unroll
		for (int r = 0; r < 9; ++r) nxt[r] = lane + 32 * r < words ? __ldcs(src + lane + 32 * r) : 0u;
		a_n = 0;
		if (lane < cnt) a_n = dirs ? rb_action_of(faces[base + lane], dirs[base + lane]) : rb_clamp_action(faces[base + lane]);
	};
	int64_t g = (int64_t)blockIdx.x * (kThreads / 32) + wib;
	if (g < n_groups) fetch(g);
	for (; g < n_groups; g += stride) {
		const int64_t base = g * kMrStates;
		const int words = (int)min((int64_t)kMrStates, n - base) * 72;
		uint32_t* dst = reinterpret_cast<uint32_t*>(out) + base * 72;
		const uint32_t a_l = a_n;
#pragma unroll
		for (int r = 0; r < 9; ++r) buf[lane + 32 * r] = nxt[r];
		if (g + stride < n_groups) fetch(g + stride);
		__syncwarp();
#pragma unroll
		for (int r = 0; r < 9; ++r) {
			const int i = lane + 32 * r;
			const uint32_t st = ((uint32_t)i * 911u) >> 16;            // i / 72 for i < 288
			const uint32_t a = __shfl_sync(0xffffffffu, a_l, st);
			if (i < words) __stcs(dst + i, gather_word(s_tab, reinterpret_cast<const uint16_t*>(buf + st * 72u), a, i - 72u * st));
		}
		__syncwarp();
	}
}

// multi_is_solved: 288 B in + 1 B out per state.  A warp takes 4 states = 288 words = 9 full coalesced rounds; each lane
// collects a 4-bit "differs from solved" mask for the states its words belong to, one warp OR-reduction finishes them.
__global__ void __launch_bounds__(kThreads)
k_is_solved(const int8_t* __restrict__ in, uint8_t* __restrict__ flags, int64_t n) {
	__shared__ uint32_t s_solved[72];
	if (threadIdx.x < 72) s_solved[threadIdx.x] = reinterpret_cast<const uint32_t*>(g_solved686)[threadIdx.x];
	__syncthreads();
	const int lane = threadIdx.x & 31;
	const int64_t n_groups = (n + 3) / 4;
	for (int64_t g = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); g < n_groups; g += (int64_t)gridDim.x * (kThreads / 32)) {
		const int64_t base = g * 4;
		const int cnt = (int)min((int64_t)4, n - base), words = cnt * 72;
		const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + base * 72;
		uint32_t v[9];
#pragma unroll
		for (int r = 0; r < 9; ++r) v[r] = lane + 32 * r < words ? __ldcs(src + lane + 32 * r) : 0u;
		uint32_t bad = 0;
#pragma unroll
		for (int r = 0; r < 9; ++r) {
			const uint32_t i = lane + 32 * r, st = (i * 911u) >> 16;          // i / 72 for i < 288
			if ((int)i < words && v[r] != s_solved[i - 72u * st]) bad |= 1u << st;
		}
		bad = __reduce_or_sync(0xffffffffu, bad);
		if (lane < cnt) flags[base + lane] = ((bad >> lane) & 1u) ^ 1u;
	}
}

// Four int8 values (one word) widened to the one-hot element type and stored at element offset 4 * w of `base`:
// f32 -> one 16-byte store, bf16 (raw bits, every int8 value is exact in bf16) -> one 8-byte store.
typedef uint16_t rb_bf16;
__device__ __forceinline__ void store_widened(float* __restrict__ base, int64_t w, uint32_t x, int pol) {
	float4 o;
	o.x = (float)(int8_t)(x & 0xff);
	o.y = (float)(int8_t)((x >> 8) & 0xff);
	o.z = (float)(int8_t)((x >> 16) & 0xff);
	o.w = (float)(int8_t)(x >> 24);
	rb_st_stream(reinterpret_cast<float4*>(base) + w, o, pol);
}
__device__ __forceinline__ void store_widened(rb_bf16* __restrict__ base, int64_t w, uint32_t x, int pol) {
	const uint32_t b0 = __float_as_uint((float)(int8_t)(x & 0xff)) >> 16, b1 = __float_as_uint((float)(int8_t)((x >> 8) & 0xff)) & 0xffff0000u;
	const uint32_t b2 = __float_as_uint((float)(int8_t)((x >> 16) & 0xff)) >> 16, b3 = __float_as_uint((float)(int8_t)(x >> 24)) & 0xffff0000u;
	uint2* p = reinterpret_cast<uint2*>(base) + w;
	const uint2 o = make_uint2(b0 | b1, b2 | b3);
	if (pol == RB_STORE_CS) __stcs(p, o);
	else *p = o;
}

// as_oh: int8 -> f32 widening of the already one-hot state: 288 B in, 1152 B out.  Thread = 4 bytes -> float4, eight
// independent loads in flight per thread (read latency under a saturating write stream is several microseconds).
template <typename OH>
__global__ void __launch_bounds__(kThreads)
k_as_oh(const int8_t* __restrict__ in, OH* __restrict__ oh, int64_t n_words, int pol) {
	const uint32_t* src = reinterpret_cast<const uint32_t*>(in);
	constexpr int kU = 8;
	const int64_t step = (int64_t)gridDim.x * blockDim.x * kU;
	for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x * kU + threadIdx.x; i0 < n_words; i0 += step) {
		uint32_t w[kU];
#pragma unroll
		for (int k = 0; k < kU; ++k) w[k] = i0 + k * kThreads < n_words ? __ldcs(src + i0 + k * kThreads) : 0u;
#pragma unroll
		for (int k = 0; k < kU; ++k) {
			if (i0 + k * kThreads < n_words) store_widened(oh, i0 + k * kThreads, w[k], pol);
		}
	}
}

// as_correct (cube.py:371-380): f32 [n][288] -> f32 [n][48]: +1 where the sticker's 6 channels equal the solved
// sticker's, else -1.  Thread = one sticker.
__global__ void __launch_bounds__(kThreads)
k_as_correct(const float* __restrict__ oh, float* __restrict__ out, int64_t n_slots) {
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_slots; i += (int64_t)gridDim.x * blockDim.x) {
		const int slot = (int)(i % kSlots);
		const float2* p = reinterpret_cast<const float2*>(oh + i * 6);
		bool ok = true;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const float2 q = p[k];
			ok &= (q.x == (float)g_solved686[slot * 6 + 2 * k]) & (q.y == (float)g_solved686[slot * 6 + 2 * k + 1]);
		}
		out[i] = ok ? 1.f : -1.f;
	}
}

// Emit one state held in shared memory (288 B at `s`): raw bytes, f32 one-hot row and solved flag; warp-wide.
template <typename OH>
__device__ __forceinline__ void warp_emit(const uint8_t* s, int lane, int64_t row, int8_t* __restrict__ states,
                                          OH* __restrict__ oh, uint8_t* __restrict__ solved, int pol) {
	if (states && lane < 18)
		rb_st_stream(reinterpret_cast<uint4*>(states + row * kStateBytes) + lane, reinterpret_cast<const uint4*>(s)[lane], pol);
	if (oh) {
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const int w = lane + 32 * k;
			if (w < 72) store_widened(oh + row * kStateBytes, w, reinterpret_cast<const uint32_t*>(s)[w], pol);
		}
	}
	if (solved) {
		bool ok = true;
#pragma unroll
		for (int k = 0; k < 3; ++k) {
			const int w = lane + 32 * k;
			if (w < 72) ok &= reinterpret_cast<const uint32_t*>(s)[w] == reinterpret_cast<const uint32_t*>(g_solved686)[w];
		}
		ok = __all_sync(0xffffffffu, ok);
		if (lane == 0) solved[row] = ok;
	}
}

// dst = move a applied to src (both 288 B in shared memory, distinct buffers); warp-wide.
__device__ __forceinline__ void warp_move(uint8_t* dst, const uint8_t* src, const uint8_t* s_perm, uint32_t a, int lane) {
	for (int slot = lane; slot < kSlots; slot += 32)
		copy_record(dst + slot * 6, src + s_perm[a * kSlots + slot] * 6);
	__syncwarp();
}

constexpr int kWarps = kThreads / 32;

// expand12 (+ one-hot + solved): warp per parent; parent and one child buffer per warp in shared memory.
template <typename OH>
__device__ __forceinline__ void warp_expand12(uint8_t* child, const uint8_t* parent, const uint8_t* s_perm, int lane,
                                              int64_t prow, int8_t* __restrict__ children, OH* __restrict__ children_oh,
                                              uint8_t* __restrict__ solved, int pol) {
	for (uint32_t a = 0; a < 12; ++a) {
		warp_move(child, parent, s_perm, a, lane);
		warp_emit(child, lane, prow * 12 + a, children, children_oh, solved, pol);
		__syncwarp();
	}
}

template <typename OH>
__global__ void __launch_bounds__(kThreads)
k_expand12(const int8_t* __restrict__ in, int8_t* __restrict__ children, OH* __restrict__ children_oh,
           uint8_t* __restrict__ solved, int64_t n, int pol) {
	__shared__ __align__(16) uint8_t s_perm[12 * 48];
	__shared__ __align__(16) uint8_t s_buf[kWarps][2][kStateBytes];
	stage_perm(s_perm);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
	const int64_t n_warps = (int64_t)gridDim.x * kWarps;
	for (int64_t i = warp; i < n; i += n_warps) {
		if (lane < 18)
			reinterpret_cast<uint4*>(s_buf[wib][0])[lane] = rb_ld_stream(reinterpret_cast<const uint4*>(in + i * kStateBytes) + lane);
		__syncwarp();
		warp_expand12(s_buf[wib][1], s_buf[wib][0], s_perm, lane, i, children, children_oh, solved, pol);
	}
}

// expand12, children only (288 B in, 12 x 288 B out per parent): the 12 children of a parent are 864 contiguous words =
// 27 full rounds of a warp; word i is word i % 72 of action i / 72, gathered from the parent held once in shared memory.
__global__ void __launch_bounds__(kThreads)
k_expand12_states(const int8_t* __restrict__ in, int8_t* __restrict__ children, int64_t n) {
	__shared__ uint16_t s_tab[12 * 72];
	__shared__ __align__(16) uint32_t s_buf[kWarps][72];
	stage_gather(s_tab);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	uint32_t* buf = s_buf[wib];
	// Word i = lane + 32 r of the 864 output words is word i % 72 of action i / 72, i.e. table entry i: the same 27 entries for
	// every parent this lane works on.  They are read once into registers (two per register); the kernel was bound by its
	// shared-memory wavefronts (l1tex 97 %), a fifth of which were these table reads.
	uint32_t pairs[14];
#pragma unroll
	for (int r = 0; r < 14; ++r) pairs[r] = (uint32_t)s_tab[lane + 64 * r] | (r < 13 ? (uint32_t)s_tab[lane + 64 * r + 32] << 16 : 0u);
	const uint16_t* src16 = reinterpret_cast<const uint16_t*>(buf);
	for (int64_t p = (int64_t)blockIdx.x * kWarps + wib; p < n; p += (int64_t)gridDim.x * kWarps) {
		const uint32_t* src = reinterpret_cast<const uint32_t*>(in) + p * 72;
		uint32_t* dst = reinterpret_cast<uint32_t*>(children) + p * 864;
		buf[lane] = __ldcs(src + lane);
		buf[lane + 32] = __ldcs(src + lane + 32);
		if (lane < 8) buf[lane + 64] = __ldcs(src + lane + 64);
		__syncwarp();
#pragma unroll
		for (int r = 0; r < 27; ++r) {
			const uint32_t pair = (r & 1) ? pairs[r >> 1] >> 16 : pairs[r >> 1] & 0xffffu;
			__stcs(dst + lane + 32 * r, (uint32_t)src16[pair & 0xffu] | ((uint32_t)src16[pair >> 8] << 16));
		}
		__syncwarp();
	}
}

// scramble: `depth` moves per cube, final state only; warp per cube, ping-pong buffers in shared memory.
__global__ void __launch_bounds__(kThreads)
k_scramble(const uint8_t* __restrict__ actions, int64_t stride_cube, int64_t stride_move,
           const int8_t* __restrict__ start, int8_t* __restrict__ out, int64_t n, int depth) {
	__shared__ __align__(16) uint8_t s_perm[12 * 48];
	__shared__ __align__(16) uint8_t s_buf[kWarps][2][kStateBytes];
	stage_perm(s_perm);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
	const int64_t n_warps = (int64_t)gridDim.x * kWarps;
	for (int64_t i = warp; i < n; i += n_warps) {
		const uint8_t* src = start ? reinterpret_cast<const uint8_t*>(start) + i * kStateBytes : g_solved686;
		if (lane < 18) reinterpret_cast<uint4*>(s_buf[wib][0])[lane] = reinterpret_cast<const uint4*>(src)[lane];
		__syncwarp();
		int cur = 0;
		for (int m0 = 0; m0 < depth; m0 += 32) {
			const int m = m0 + lane;
			const uint32_t a_l = m < depth ? rb_clamp_action(actions[i * stride_cube + (int64_t)m * stride_move]) : 0u;
			const int stop = min(32, depth - m0);
			for (int k = 0; k < stop; ++k) {
				warp_move(s_buf[wib][cur ^ 1], s_buf[wib][cur], s_perm, __shfl_sync(0xffffffffu, a_l, k), lane);
				cur ^= 1;
			}
		}
		if (lane < 18)
			rb_st_stream(reinterpret_cast<uint4*>(out + i * kStateBytes) + lane, reinterpret_cast<const uint4*>(s_buf[wib][cur])[lane], RB_STORE_CS);
		__syncwarp();
	}
}

// Render: 6x8x6 state of a move sequence from the 20x24 state of the SAME sequence.  Moves only permute sticker records,
// so "start, then the sequence" = start gathered through the sequence's net permutation, and that permutation is what
// the 20x24 state encodes: cubie c with value v carries its sticker k from its home slot to slot dst[c][v][k] (tables
// derived in rb_tables.cuh by running both representations side by side).  This lets the 6x8x6 scramble run on the
// slot-major macro-move kernel (rb_scramble_macro.cuh) instead of moving 288 bytes through shared memory per move.
// In place: the 20x24 state is parked in the first 20 bytes of each 288-byte output row.  Warp per cube.
constexpr int kStickerDst = 20 * 24 * 3, kStickerBytes = kStickerDst + 20 * 3 + 4;
__global__ void __launch_bounds__(kThreads)
k_render_from2024(int8_t* __restrict__ io, const int8_t* __restrict__ start, int64_t n, const int8_t* __restrict__ parked = nullptr) {
	// parked != nullptr: the 20x24 states come from their own int8 [n][20] array instead of the heads of the output rows
	__shared__ __align__(16) uint8_t s_tab[kStickerBytes];
	__shared__ __align__(16) uint8_t s_out[kWarps][kStateBytes];
	__shared__ __align__(16) uint8_t s_src[kWarps][kStateBytes];
	for (int i = threadIdx.x; i < kStickerBytes / 4; i += blockDim.x)
		reinterpret_cast<uint32_t*>(s_tab)[i] = reinterpret_cast<const uint32_t*>(g_stickers686)[i];
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int64_t stride = (int64_t)gridDim.x * kWarps;
	// Per-lane constants of its two stickers j = lane, lane + 32 (24 corner stickers, then 24 edge stickers): owning cubie,
	// table row base, home slot (source record) and, for a solved start, the record itself (one-hot of the home face's colour).
	int cub[2], tab[2], srcoff[2];
	uint16_t solved_rec[2][3];
#pragma unroll
	for (int r = 0; r < 2; ++r) {
		const int j = lane + 32 * r;
		const int c = j < 24 ? j / 3 : 8 + ((j - 24) >> 1), k = j < 24 ? j - 3 * (j / 3) : (j - 24) & 1;
		cub[r] = c < 20 ? c : 0;
		tab[r] = c * 72 + k;
		const uint32_t src = j < 48 ? s_tab[kStickerDst + c * 3 + k] : 0u;
		srcoff[r] = (int)src * 6;
		const uint32_t face = src >> 3, bit = 1u << (8 * (face & 1u));
#pragma unroll
		for (int h = 0; h < 3; ++h) solved_rec[r][h] = (uint16_t)((face >> 1) == (uint32_t)h ? bit : 0u);
	}
	// Work unit = 32 consecutive cubes per warp.  Read latency under a saturating write stream is several microseconds, so the
	// parked 20x24 states of the warp's NEXT 32 cubes (lane l: the 20 bytes of cube l) are fetched into registers while the
	// current 32 are rendered from shared memory; the start row (if any) is fetched one cube ahead.
	__shared__ __align__(16) uint32_t s_park[kWarps][32 * 5];
	const int64_t n_chunks = (n + 31) / 32;
	int64_t ch = (int64_t)blockIdx.x * kWarps + wib;
	uint32_t pk[5] = {0u, 0u, 0u, 0u, 0u};
	auto fetch_parked = [&](int64_t c) {
		const int64_t cube = c * 32 + lane;
		if (c < n_chunks && cube < n) {
			if (parked) {
				const uint32_t* q = reinterpret_cast<const uint32_t*>(parked + cube * 20);
#pragma unroll
				for (int k = 0; k < 5; ++k) pk[k] = q[k];
			} else {
				const uint4 q = *reinterpret_cast<const uint4*>(io + cube * kStateBytes);
				pk[0] = q.x; pk[1] = q.y; pk[2] = q.z; pk[3] = q.w;
				pk[4] = *reinterpret_cast<const uint32_t*>(io + cube * kStateBytes + 16);
			}
		}
	};
	uint4 src_next = make_uint4(0u, 0u, 0u, 0u);
	auto fetch_start = [&](int64_t cube) {
		if (start && cube < n && lane < 18) src_next = rb_ld_stream(reinterpret_cast<const uint4*>(start + cube * kStateBytes) + lane);
	};
	fetch_parked(ch);
	for (; ch < n_chunks; ch += stride) {
		const int64_t base = ch * 32;
		const int cnt = (int)min((int64_t)32, n - base);
#pragma unroll
		for (int k = 0; k < 5; ++k) s_park[wib][lane * 5 + k] = pk[k];
		fetch_parked(ch + stride);
		fetch_start(base);
		__syncwarp();
		for (int t = 0; t < cnt; ++t) {
			int8_t* row = io + (base + t) * kStateBytes;
			const uint32_t v = lane < 20 ? (uint32_t)reinterpret_cast<const uint8_t*>(s_park[wib])[t * 20 + lane] : 0u;
			if (start && lane < 18) reinterpret_cast<uint4*>(s_src[wib])[lane] = src_next;
			if (t + 1 < cnt) fetch_start(base + t + 1);
			__syncwarp();
#pragma unroll
			for (int r = 0; r < 2; ++r) {
				uint32_t val = __shfl_sync(0xffffffffu, v, cub[r]);
				if (lane + 32 * r < 48) {
					val = val < 24u ? val : 0u;
					const uint32_t dst = s_tab[tab[r] + val * 3];
					uint16_t h0 = solved_rec[r][0], h1 = solved_rec[r][1], h2 = solved_rec[r][2];
					if (start) {
						const uint16_t* q = reinterpret_cast<const uint16_t*>(s_src[wib] + srcoff[r]);
						h0 = q[0]; h1 = q[1]; h2 = q[2];
					}
					uint16_t* o = reinterpret_cast<uint16_t*>(s_out[wib] + dst * 6);
					o[0] = h0; o[1] = h1; o[2] = h2;
				}
			}
			__syncwarp();
			if (lane < 18) rb_st_stream(reinterpret_cast<uint4*>(row) + lane, reinterpret_cast<const uint4*>(s_out[wib])[lane], RB_STORE_CS);
			__syncwarp();
		}
	}
}

// 6x8x6 state -> the 20x24 state of the same cube: cubie c has value v iff every sticker k of c shows its home colour in slot
// dst[c][v][k] (the sticker tables of rb_tables.cuh read the other way round).  Thread per (state, cubie); value 255 and
// ok = 0 when no value fits (not a reachable cube).  Used to run searches on the 20-byte representation whatever the caller's.
__global__ void __launch_bounds__(kThreads)
k_as2024(const int8_t* __restrict__ states, int8_t* __restrict__ out, uint8_t* __restrict__ ok, int64_t n) {
	__shared__ __align__(16) uint8_t s_tab[kStickerBytes];
	for (int i = threadIdx.x; i < kStickerBytes / 4; i += blockDim.x)
		reinterpret_cast<uint32_t*>(s_tab)[i] = reinterpret_cast<const uint32_t*>(g_stickers686)[i];
	__syncthreads();
	const int64_t stride = (int64_t)gridDim.x * blockDim.x;
	for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n * 20; t += stride) {
		const int64_t i = t / 20;
		const int c = (int)(t - i * 20), K = c < 8 ? 3 : 2;
		const int8_t* st = states + i * kStateBytes;
		int found = 255;
		for (int v = 0; v < 24; ++v) {
			bool fits = true;
			for (int k = 0; k < K; ++k) {
				const int slot = s_tab[(c * 24 + v) * 3 + k], colour = s_tab[kStickerDst + c * 3 + k] >> 3;
				fits = fits && st[slot * 6 + colour] == 1;
			}
			if (fits && found == 255) found = v;
		}
		out[i * 20 + c] = (int8_t)found;
		if (found == 255 && ok) ok[i] = 0;
	}
}

// sequence_scramble / fused ADI generator: same unit decomposition as the 20x24 kernel (game, chunk of depth).
template <bool kChildren, typename OH>
__global__ void __launch_bounds__(kThreads)
k_sequence(const uint8_t* __restrict__ faces, const uint8_t* __restrict__ dirs, int games, int depth,
           int with_solved, int chunk, int8_t* __restrict__ states, OH* __restrict__ oh,
           uint8_t* __restrict__ solved_states, int8_t* __restrict__ children, OH* __restrict__ children_oh,
           uint8_t* __restrict__ solved_children, int pol) {
	__shared__ __align__(16) uint8_t s_perm[12 * 48];
	__shared__ __align__(16) uint8_t s_buf[kWarps][3][kStateBytes];
	stage_perm(s_perm);
	__syncthreads();
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
	const int chunks_per_game = (depth + chunk - 1) / chunk;
	const int64_t n_units = (int64_t)games * chunks_per_game;
	const int64_t warp = (int64_t)blockIdx.x * kWarps + wib;
	const int64_t n_warps = (int64_t)gridDim.x * kWarps;
	for (int64_t u = warp; u < n_units; u += n_warps) {
		const int g = (int)(u / chunks_per_game);
		const int d0 = (int)(u % chunks_per_game) * chunk;
		const int d1 = min(depth, d0 + chunk);
		if (lane < 18) reinterpret_cast<uint4*>(s_buf[wib][0])[lane] = reinterpret_cast<const uint4*>(g_solved686)[lane];
		__syncwarp();
		int cur = 0;
		const int total = d1 - with_solved;
		int applied = 0, buf0 = 0;
		uint32_t a_l = 0;
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
				warp_move(s_buf[wib][cur ^ 1], s_buf[wib][cur], s_perm, __shfl_sync(0xffffffffu, a_l, applied - buf0), lane);
				cur ^= 1;
				++applied;
			}
			const int64_t row = (int64_t)g * depth + d;
			warp_emit(s_buf[wib][cur], lane, row, states, oh, solved_states, pol);
			if (kChildren) warp_expand12(s_buf[wib][2], s_buf[wib][cur], s_perm, lane, row, children, children_oh, solved_children, pol);
			__syncwarp();
		}
	}
}

}  // namespace rb686
