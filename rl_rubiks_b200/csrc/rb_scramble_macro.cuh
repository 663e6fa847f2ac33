// Fast multi-move scramble for the 20x24 representation (BASELINE configs[1]: 2^24 cubes x 100 moves).
//
// The byte-LUT kernel (rb2024::k_scramble) does 20 data-dependent shared-memory lookups per move and is bound by
// shared-memory bank conflicts at ~2 % of the HBM roofline (profiles/r1a_scramble_v1_ncu.txt).  This kernel changes the
// representation instead of the table: inside the kernel a cube is kept "slot-major" -- which cubie sits in each of
// the 8 corner / 12 edge positions -- so that a move is a FIXED byte permutation of registers (PRMT with a selector
// that depends only on the action) plus an additive orientation update:
//   corners: 8 id bytes (C0,C1) + 8 twist accumulators (W0,W1); twist in Z3, move adds a per-slot constant
//   edges  : 12 bytes id | flip<<4 (E0,E1,E2); move XORs a per-slot constant into the flip bit
// Permutation + additive orientation is closed under composition, so three consecutive moves are fused into one
// table row (13^3 rows, index a0 + 13 a1 + 169 a2, action 12 = identity padding): 24 B per row, 52.7 KB in shared memory.
// Per 3 moves a thread does 3 LDS.64 + 10 PRMT + ~14 other ALU ops instead of 60 LDS.U8 + ~180 ALU ops.
// The reference's cubie-major int8[20] state is rebuilt once per cube at the end (scatter through shared memory).
//
// The corner twist t relates to the reference's orientation o (the axis the tracked sticker faces, maps.py:128)
// by t = o for positions {1,3,4,6} and t = -o mod 3 for positions {0,2,5,7} (the corner's chirality, the same split
// cube.py:292 uses); with that labelling every quarter turn adds a constant per slot.  rbs::build() derives all
// rows from the 20x24 LUT and verifies the additivity for every (action, position, orientation).
//
// Action tiles ([T cubes][depth] bytes, contiguous in HBM) are brought into shared memory with one 1-D bulk
// async copy (cp.async.bulk, TMA engine, completion on an mbarrier), double buffered against the compute.
#pragma once
#include "rb_common.cuh"
#include "rb_tables.cuh"
#include <stdlib.h>

namespace rbs {

constexpr int kA = 13;                          // 12 actions + identity
constexpr int kRows = kA * kA * kA;             // 2197
constexpr int kRowWords = 6;                    // s_corner, s_edge0, s_edge1, s_edge2, twists, flips
constexpr int kTableBytes = (kRows * kRowWords * 4 + 15) / 16 * 16;   // 52,736
constexpr int kMaxThreads = 768;
constexpr int kSmemBudget = 227 * 1024;

__device__ __align__(16) uint32_t g_macro[kTableBytes / 4];

struct Elem {                                   // one cube-group element in slot-major gather form
	uint8_t csrc[8], ctw[8], esrc[12], efl[12];
};

static inline bool neg_chirality(int pos) { return pos == 0 || pos == 2 || pos == 5 || pos == 7; }
static inline int twist_of(int pos, int ori) { return neg_chirality(pos) ? (3 - ori) % 3 : ori; }

// Single move as a gather element, derived from the direct LUT (and checked for additivity).
static bool single(const rbt::Tables& t, int a, Elem& e) {
	for (int q = 0; q < 8; ++q) { e.csrc[q] = (uint8_t)q; e.ctw[q] = 0; }
	for (int q = 0; q < 12; ++q) { e.esrc[q] = (uint8_t)q; e.efl[q] = 0; }
	if (a == 12) return true;
	for (int p = 0; p < 8; ++p) {
		int delta = -1, dst = -1;
		for (int o = 0; o < 3; ++o) {
			const int v = t.lut[a][0][3 * p + o], p2 = v / 3, o2 = v % 3;
			const int d = (twist_of(p2, o2) - twist_of(p, o) + 3) % 3;
			if (o == 0) { delta = d; dst = p2; }
			else if (d != delta || p2 != dst) return false;
		}
		e.csrc[dst] = (uint8_t)p;
		e.ctw[dst] = (uint8_t)delta;
	}
	for (int p = 0; p < 12; ++p) {
		const int v0 = t.lut[a][1][2 * p], v1 = t.lut[a][1][2 * p + 1];
		if (v0 / 2 != v1 / 2 || ((v0 % 2) ^ 0) != ((v1 % 2) ^ 1)) return false;
		e.esrc[v0 / 2] = (uint8_t)p;
		e.efl[v0 / 2] = (uint8_t)(v0 % 2);
	}
	return true;
}

// r = "first x, then y"
static void compose(const Elem& x, const Elem& y, Elem& r) {
	for (int q = 0; q < 8; ++q) {
		r.csrc[q] = x.csrc[y.csrc[q]];
		r.ctw[q] = (uint8_t)((x.ctw[y.csrc[q]] + y.ctw[q]) % 3);
	}
	for (int q = 0; q < 12; ++q) {
		r.esrc[q] = x.esrc[y.esrc[q]];
		r.efl[q] = (uint8_t)(x.efl[y.esrc[q]] ^ y.efl[q]);
	}
}

// Register layout: C0/W0 = corner slots 0-3 (byte i = slot i), C1/W1 = slots 4-7; E0,E1,E2 = edge slots 0-3, 4-7, 8-11.
static void encode(const Elem& e, uint32_t* row) {
	uint32_t sc = 0, tw = 0, fl = 0;
	for (int q = 0; q < 8; ++q) sc |= (uint32_t)e.csrc[q] << (4 * q);       // low 16: selector of C0', high 16: of C1'
	for (int i = 0; i < 4; ++i) {
		tw |= (uint32_t)e.ctw[i] << (8 * i);                                // W0 += tw & 0x0f0f0f0f
		tw |= (uint32_t)e.ctw[4 + i] << (8 * i + 4);                        // W1 += (tw >> 4) & 0x0f0f0f0f
	}
	row[0] = sc;
	for (int d = 0; d < 3; ++d) {
		uint32_t sa = 0, sb = 0;
		for (int i = 0; i < 4; ++i) {
			const int s = e.esrc[4 * d + i];
			sa |= (uint32_t)(s < 8 ? s : 0) << (4 * i);                     // x  = prmt(E0, E1, sa)
			sb |= (uint32_t)(s < 8 ? i : 4 + (s - 8)) << (4 * i);           // Ed' = prmt(x, E2, sb)
			fl |= (uint32_t)e.efl[4 * d + i] << (8 * i + 4 + d);            // Ed' ^= (fl >> d) & 0x10101010
		}
		row[1 + d] = sa | (sb << 16);
	}
	row[4] = tw;
	row[5] = fl;
}

struct Host {
	uint32_t rows[kTableBytes / 4];
	bool ok;
};

static const Host& host() {
	static Host h;
	static std::once_flag once;
	std::call_once(once, [] {
		memset(&h, 0, sizeof(h));
		const rbt::Tables& t = rbt::host();
		Elem s[kA];
		h.ok = true;
		for (int a = 0; a < kA; ++a) h.ok = single(t, a, s[a]) && h.ok;
		for (int a2 = 0; a2 < kA; ++a2)
			for (int a1 = 0; a1 < kA; ++a1)
				for (int a0 = 0; a0 < kA; ++a0) {
					Elem x, y;
					compose(s[a0], s[a1], x);
					compose(x, s[a2], y);
					encode(y, h.rows + (a0 + kA * a1 + kA * kA * a2) * kRowWords);
				}
	});
	return h;
}

// ---- device side ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
	asm volatile(
		"{\n\t"
		".reg .pred p;\n\t"
		"WAIT_%=:\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
		"@p bra DONE_%=;\n\t"
		"bra WAIT_%=;\n\t"
		"DONE_%=:\n\t"
		"}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk async copy global -> shared (TMA engine; SASS UBLKCP), completion counted in bytes on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
	             "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}

struct Slots {
	uint32_t C0, C1, W0, W1, E0, E1, E2;
};

// PRMT straight from PTX: __byte_perm() masks the selector with 0x7777 first (one extra LOP3 per permute); the table's
// selector nibbles are always < 8, so the mask is dead weight on the ALU pipe that bounds this kernel.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
	return r;
}
// x >> k on the FMA pipe (IMAD.HI by 2^(32-k)) instead of SHF on the ALU pipe.
template <int k>
__device__ __forceinline__ uint32_t shr_fma(uint32_t x) {
	uint32_t r;
	asm("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(1u << (32 - k)));
	return r;
}

// One table row applied to the slot-major state.  p1 = {corner selectors, edge selectors 0..2}, p2 = {twists, flips}.
__device__ __forceinline__ void apply_row(const uint4 p1, const uint2 p2, Slots& s) {
	const uint32_t sc_hi = shr_fma<16>(p1.x);
	const uint32_t c0 = prmt(s.C0, s.C1, p1.x), c1 = prmt(s.C0, s.C1, sc_hi);
	const uint32_t w0 = prmt(s.W0, s.W1, p1.x), w1 = prmt(s.W0, s.W1, sc_hi);
	s.C0 = c0;
	s.C1 = c1;
	s.W0 = w0 + (p2.x & 0x0f0f0f0fu);
	s.W1 = w1 + (shr_fma<4>(p2.x) & 0x0f0f0f0fu);
	const uint32_t x0 = prmt(s.E0, s.E1, p1.y), x1 = prmt(s.E0, s.E1, p1.z), x2 = prmt(s.E0, s.E1, p1.w);
	const uint32_t e0 = prmt(x0, s.E2, shr_fma<16>(p1.y)), e1 = prmt(x1, s.E2, shr_fma<16>(p1.z)), e2 = prmt(x2, s.E2, shr_fma<16>(p1.w));
	s.E0 = e0 ^ (p2.y & 0x10101010u);
	s.E1 = e1 ^ (shr_fma<1>(p2.y) & 0x10101010u);
	s.E2 = e2 ^ (shr_fma<2>(p2.y) & 0x10101010u);
}

// Table placement in shared memory.  Rows are split into a 16-byte part (selectors) and an 8-byte part (twists, flips).
//   kMoves == 3: 13^3 rows, one copy: 52.7 KB; random rows => ~2.5-way bank conflicts on every load.
//   kMoves == 2: 13^2 rows, the 16-byte part replicated 8x and the 8-byte part 16x so that lane l always reads bank
//                group l%8 (resp. bank pair l%16): every LDS.128 / LDS.64 is conflict free (43 KB).
template <int kMoves>
struct Layout;
template <>
struct Layout<3> {
	static constexpr int kRowsL = kA * kA * kA, kP1 = kRowsL * 16, kBytes = (kRowsL * 24 + 15) / 16 * 16;
	static __device__ __forceinline__ void apply(const uint8_t* table, uint32_t idx, int, Slots& s) {
		idx = min(idx, (uint32_t)(kRowsL - 1));
		apply_row(*reinterpret_cast<const uint4*>(table + idx * 16u), *reinterpret_cast<const uint2*>(table + kP1 + idx * 8u), s);
	}
	static __device__ __forceinline__ void fill(uint8_t* table, const uint32_t* g_rows) {
		for (int i = threadIdx.x; i < kRowsL; i += blockDim.x) {
			const uint32_t* r = g_rows + i * kRowWords;
			*reinterpret_cast<uint4*>(table + i * 16) = make_uint4(r[0], r[1], r[2], r[3]);
			*reinterpret_cast<uint2*>(table + kP1 + i * 8) = make_uint2(r[4], r[5]);
		}
	}
};
template <>
struct Layout<2> {
	static constexpr int kRowsL = kA * kA, kP1 = kRowsL * 8 * 16, kBytes = kP1 + kRowsL * 16 * 8;
	static __device__ __forceinline__ void apply(const uint8_t* table, uint32_t idx, int lane, Slots& s) {
		idx = min(idx, (uint32_t)(kRowsL - 1));
		// both offsets as one IMAD each (FMA pipe); the compiler's shift + OR would sit on the ALU pipe
		uint32_t o1, o2;
		asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(o1) : "r"(idx), "r"((uint32_t)(lane & 7) * 16u));
		asm("mad.lo.u32 %0, %1, 128, %2;" : "=r"(o2) : "r"(idx), "r"((uint32_t)(lane & 15) * 8u + (uint32_t)kP1));
		apply_row(*reinterpret_cast<const uint4*>(table + o1), *reinterpret_cast<const uint2*>(table + o2), s);
	}
	static __device__ __forceinline__ void fill(uint8_t* table, const uint32_t* g_rows) {
		for (int i = threadIdx.x; i < kRowsL * 16; i += blockDim.x) {
			const int row = i >> 4, c = i & 15;
			const uint32_t* r = g_rows + (row + kA * kA * 12) * kRowWords;      // a0 + 13 a1, third move = identity
			if (c < 8) *reinterpret_cast<uint4*>(table + (row * 8 + c) * 16) = make_uint4(r[0], r[1], r[2], r[3]);
			*reinterpret_cast<uint2*>(table + kP1 + (row * 16 + c) * 8) = make_uint2(r[4], r[5]);
		}
	}
};

__device__ __forceinline__ uint32_t mod3_bytes(uint32_t w) {
	uint32_t r = 0;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		const uint32_t t = (w >> (8 * i)) & 0xffu;
		r |= (t - 3u * ((t * 171u) >> 9)) << (8 * i);
	}
	return r;
}

// Slot-major -> the reference's cubie-major int8[20] (written to this thread's 20-byte row in shared memory).
__device__ __forceinline__ void store_state(uint8_t* __restrict__ o, const Slots& s) {
	const uint32_t w0 = mod3_bytes(s.W0), w1 = mod3_bytes(s.W1);
#pragma unroll
	for (int q = 0; q < 8; ++q) {
		const uint32_t id = ((q < 4 ? s.C0 : s.C1) >> (8 * (q & 3))) & 7u;
		const uint32_t t = ((q < 4 ? w0 : w1) >> (8 * (q & 3))) & 3u;
		const bool neg = q == 0 || q == 2 || q == 5 || q == 7;
		const uint32_t ori = neg ? (t ? 3u - t : 0u) : t;
		o[id] = (uint8_t)(3 * q + ori);
	}
#pragma unroll
	for (int q = 0; q < 12; ++q) {
		const uint32_t b = ((q < 4 ? s.E0 : (q < 8 ? s.E1 : s.E2)) >> (8 * (q & 3))) & 0xffu;
		o[8 + (b & 15u)] = (uint8_t)(2 * q + ((b >> 4) & 1u));
	}
}

// Dynamic shared memory: [macro table | mbarrier | action tile T*depth].  One action buffer per CTA and two CTAs per SM:
// while one CTA waits for its bulk copy the other computes, and 40 warps per SM keep both math pipes fed.  The finished
// state overwrites the first 20 bytes of the thread's own (consumed) action row, so no separate output tile is needed.
//   kDouble = true : one CTA per SM, two action buffers, tile j+1 is copied while tile j is computed
//   kDouble = false: two CTAs per SM with one buffer each, the other CTA computes while this one waits for its copy
template <int kMoves, bool kWordAligned, bool kDouble>
__global__ void __launch_bounds__(kMaxThreads, kDouble ? 1 : 2)
k_scramble_macro(const uint8_t* __restrict__ actions, int8_t* __restrict__ out, int64_t n, int depth, int buf_bytes) {
	using L = Layout<kMoves>;
	extern __shared__ __align__(128) uint8_t smem[];
	uint8_t* table = smem;
	uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBytes);
	uint8_t* buf0 = smem + L::kBytes + 64;
	const int T = blockDim.x;
	const int lane = threadIdx.x & 31;
	constexpr int kBufs = kDouble ? 2 : 1;

	const int64_t n_tiles = (n + T - 1) / T;
	const int64_t my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	const int64_t tile_bytes = (int64_t)T * depth;

	auto issue = [&](int64_t j) {               // thread 0: start the bulk copy of my j-th tile into buffer j % kBufs
		const int64_t tile = blockIdx.x + j * gridDim.x;
		const int cnt = (int)min((int64_t)T, n - tile * T);
		const uint32_t bulk = (uint32_t)(((int64_t)cnt * depth) & ~15ll);
		if (bulk) {
			mbar_expect_tx(&bars[j % kBufs], bulk);
			bulk_g2s(buf0 + (j % kBufs) * buf_bytes, actions + tile * tile_bytes, bulk, &bars[j % kBufs]);
		}
	};
	if (threadIdx.x == 0) {
		mbar_init(&bars[0], 1);
		mbar_init(&bars[1], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	if (threadIdx.x == 0 && my_tiles > 0) issue(0);
	L::fill(table, g_macro);
	__syncthreads();

	for (int64_t j = 0; j < my_tiles; ++j) {
		const int64_t tile = blockIdx.x + j * gridDim.x;
		const int cnt = (int)min((int64_t)T, n - tile * T);
		const int bytes = cnt * depth, bulk = bytes & ~15;
		uint8_t* buf = buf0 + (j % kBufs) * buf_bytes;
		if (kDouble && threadIdx.x == 0 && j + 1 < my_tiles) issue(j + 1);   // the other buffer was released by the barrier ending tile j-1
		if (bulk < bytes) {                                            // ragged last tile: < 16 trailing bytes by hand
			if (threadIdx.x < bytes - bulk) buf[bulk + threadIdx.x] = actions[tile * tile_bytes + bulk + threadIdx.x];
			__syncthreads();
		}
		if (bulk) mbar_wait(&bars[j % kBufs], (uint32_t)((j / kBufs) & 1));

		if (threadIdx.x < cnt) {
			uint8_t* row = buf + threadIdx.x * depth;
			Slots s{0x03020100u, 0x07060504u, 0u, 0u, 0x03020100u, 0x07060504u, 0x0b0a0908u};
			int m = 0, steps = 0;
			if (kMoves == 3) {
				for (; m + 12 <= depth; m += 12) {                     // 12 moves = 3 words = 4 rows
					uint32_t w0, w1, w2;
					if (kWordAligned) {
						const uint32_t* p = reinterpret_cast<const uint32_t*>(row + m);
						w0 = p[0]; w1 = p[1]; w2 = p[2];
					} else {
						const uint8_t* p = row + m;
						w0 = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
						w1 = p[4] | (p[5] << 8) | (p[6] << 16) | ((uint32_t)p[7] << 24);
						w2 = p[8] | (p[9] << 8) | (p[10] << 16) | ((uint32_t)p[11] << 24);
					}
					w0 = __vminu4(w0, 0x0c0c0c0cu); w1 = __vminu4(w1, 0x0c0c0c0cu); w2 = __vminu4(w2, 0x0c0c0c0cu);
					L::apply(table, __dp4a(w0, 0x00A90D01u, 0u), lane, s);
					L::apply(table, __dp4a(w1, 0x0000A90Du, __dp4a(w0, 0x01000000u, 0u)), lane, s);
					L::apply(table, __dp4a(w2, 0x000000A9u, __dp4a(w1, 0x0D010000u, 0u)), lane, s);
					L::apply(table, __dp4a(w2, 0xA90D0100u, 0u), lane, s);
					if (((steps += 4) & 63) == 0) { s.W0 = mod3_bytes(s.W0); s.W1 = mod3_bytes(s.W1); }   // accumulators stay < 256
				}
				for (; m < depth; m += 3) {                            // up to 11 trailing moves, identity padded
					const uint32_t a0 = row[m], a1 = m + 1 < depth ? row[m + 1] : 12u, a2 = m + 2 < depth ? row[m + 2] : 12u;
					L::apply(table, min(a0, 12u) + 13u * min(a1, 12u) + 169u * min(a2, 12u), lane, s);
				}
			} else {
#pragma unroll 2
				for (; m + 4 <= depth; m += 4) {                       // 4 moves = 1 word = 2 rows
					uint32_t w;
					if (kWordAligned) w = *reinterpret_cast<const uint32_t*>(row + m);
					else w = row[m] | (row[m + 1] << 8) | (row[m + 2] << 16) | ((uint32_t)row[m + 3] << 24);
					w = __vminu4(w, 0x0c0c0c0cu);                       // out-of-range actions become the identity: stays in the table
					L::apply(table, __dp4a(w, 0x00000D01u, 0u), lane, s);
					L::apply(table, __dp4a(w, 0x0D010000u, 0u), lane, s);
					if (((steps += 2) & 63) == 0) { s.W0 = mod3_bytes(s.W0); s.W1 = mod3_bytes(s.W1); }
				}
				for (; m < depth; m += 2) {
					const uint32_t a0 = row[m], a1 = m + 1 < depth ? row[m + 1] : 12u;
					L::apply(table, min(a0, 12u) + 13u * min(a1, 12u), lane, s);
				}
			}
			store_state(row, s);                                       // depth >= 20: the row is consumed, reuse its head
		}
		__syncthreads();
		// copy out: word k of cube c sits at buf + c*depth + 4k; consecutive threads write consecutive global words
		{
			uint8_t* dst = reinterpret_cast<uint8_t*>(out) + tile * T * 20;
			if (kWordAligned && (reinterpret_cast<uintptr_t>(out) & 3u) == 0) {
				for (int i = threadIdx.x; i < cnt * 5; i += T) {
					const int c = i / 5, k = i - 5 * c;
					reinterpret_cast<uint32_t*>(dst)[i] = *reinterpret_cast<const uint32_t*>(buf + c * depth + 4 * k);
				}
			} else {
				for (int i = threadIdx.x; i < cnt * 20; i += T) {
					const int c = i / 20, k = i - 20 * c;
					dst[i] = buf[c * depth + k];
				}
			}
		}
		__syncthreads();                                               // buffer free for the next bulk copy
		if (!kDouble && threadIdx.x == 0 && j + 1 < my_tiles) issue(j + 1);
	}
}

static int env_int(const char* name, int dflt) {
	const char* e = getenv(name);
	return e ? atoi(e) : dflt;
}
// Tuning knobs (defaults = the fastest measured on B200, profiles/r1c_*): RB_SCRAMBLE_MACRO=2|3 moves per table row,
// RB_SCRAMBLE_DOUBLE=0|1 double buffering (1 CTA/SM) vs two single-buffered CTAs per SM, RB_SCRAMBLE_THREADS tile cap.
static int macro_moves() { static int v = env_int("RB_SCRAMBLE_MACRO", 3) == 2 ? 2 : 3; return v; }
static bool double_buffered() { static bool v = env_int("RB_SCRAMBLE_DOUBLE", 1) != 0; return v; }
static int max_threads() {
	static int v = [] {
		int t = env_int("RB_SCRAMBLE_THREADS", double_buffered() ? kMaxThreads : 640);
		return t >= 32 && t <= kMaxThreads ? t / 32 * 32 : 640;
	}();
	return v;
}

static int ensure_device() {
	static std::mutex mu;
	static bool done[64] = {};
	int dev = 0;
	RB_CUDA(cudaGetDevice(&dev));
	std::lock_guard<std::mutex> lock(mu);
	if (dev < 0 || dev >= 64) return rb_fail(RB_ERR_BAD_ARG, "device ordinal out of range%s%s");
	if (done[dev]) return RB_OK;
	const Host& h = host();
	if (!h.ok) return rb_fail(RB_ERR_BAD_ARG, "macro-move table: corner twist is not additive for these move tables%s%s");
	RB_CUDA(cudaMemcpyToSymbol(g_macro, h.rows, sizeof(h.rows)));
#define RB_SET_SMEM(M, A, D) RB_CUDA(cudaFuncSetAttribute(k_scramble_macro<M, A, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget))
	RB_SET_SMEM(2, true, true); RB_SET_SMEM(2, false, true); RB_SET_SMEM(2, true, false); RB_SET_SMEM(2, false, false);
	RB_SET_SMEM(3, true, true); RB_SET_SMEM(3, false, true); RB_SET_SMEM(3, true, false); RB_SET_SMEM(3, false, false);
#undef RB_SET_SMEM
	done[dev] = true;
	return RB_OK;
}

static int64_t fixed_smem() { return (macro_moves() == 3 ? Layout<3>::kBytes : Layout<2>::kBytes) + 64; }

// Cubes per tile (= threads per block) for a given depth, 0 when the fast path does not apply.
static int tile_for(int64_t n, int depth, int* ctas_per_sm = nullptr) {
	if (depth < 20) return 0;                   // the result is written over the head of the 20+ byte action row
	int ctas = double_buffered() ? 1 : 2;
	int64_t t;
	if (ctas == 2) {
		t = ((kSmemBudget - 2048) / 2 - fixed_smem()) / depth / 32 * 32;
		if (t < 128) ctas = 1;                  // deep scrambles: one CTA per SM
	}
	if (ctas == 1) t = (kSmemBudget - 1024 - fixed_smem()) / ((double_buffered() ? 2 : 1) * (int64_t)depth + 1) / 32 * 32;
	if (t > max_threads()) t = max_threads();
	const int64_t spread = ((n + RB_NUM_SMS * ctas - 1) / (RB_NUM_SMS * ctas) + 31) / 32 * 32;      // small n: use every SM
	if (t > spread) t = spread;
	if (ctas_per_sm) *ctas_per_sm = ctas;
	return t >= 32 ? (int)t : 0;
}

template <int M, bool D>
static void launch_variant(bool al, int grid, int T, size_t smem, cudaStream_t st, const uint8_t* actions, int8_t* out, int64_t n,
                           int depth, int buf_bytes) {
	if (al) k_scramble_macro<M, true, D><<<grid, T, smem, st>>>(actions, out, n, depth, buf_bytes);
	else k_scramble_macro<M, false, D><<<grid, T, smem, st>>>(actions, out, n, depth, buf_bytes);
}

static int launch(const uint8_t* actions, int8_t* out, int64_t n, int depth, cudaStream_t st) {
	int rc = ensure_device();
	if (rc != RB_OK) return rc;
	int ctas = 1;
	const int T = tile_for(n, depth, &ctas);
	const bool dbl = double_buffered();
	const int64_t tiles = (n + T - 1) / T;
	const int64_t cap = (int64_t)RB_NUM_SMS * ctas;
	const int grid = (int)(tiles < cap ? tiles : cap);
	const int buf_bytes = (int)(((int64_t)T * depth + 15) / 16 * 16);
	const size_t smem = (size_t)fixed_smem() + (size_t)buf_bytes * (dbl ? 2 : 1);
	const bool al = depth % 4 == 0;
	if (macro_moves() == 3) {
		if (dbl) launch_variant<3, true>(al, grid, T, smem, st, actions, out, n, depth, buf_bytes);
		else launch_variant<3, false>(al, grid, T, smem, st, actions, out, n, depth, buf_bytes);
	} else {
		if (dbl) launch_variant<2, true>(al, grid, T, smem, st, actions, out, n, depth, buf_bytes);
		else launch_variant<2, false>(al, grid, T, smem, st, actions, out, n, depth, buf_bytes);
	}
	RB_LAUNCHED("scramble_macro_2024");
	return RB_OK;
}

}  // namespace rbs
