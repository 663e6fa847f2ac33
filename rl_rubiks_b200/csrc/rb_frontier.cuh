// Search-frontier kernels: open-addressing hash set on the packed state with the reference's batch-order
// index semantics, 12-neighbour frontier expansion with dedup and in-order compaction.
//
// Replaces the Python dict keyed on `state.tostring()` of BFS / AStar / MCTS
// (reference: librubiks/solving/agents.py:103-121, 286-306, 517-526, 605-609).
//
// Table layout in caller-owned HBM (capacity C slots, C a power of two, 24 B per slot):
//   keys     [C] 2 x u64   packed state; all-ones = empty.  Claimed with one 128-bit atom.cas.
//   vals     [C] i32       1-based index of the state (insertion order); 0 = inserted in the running batch
//   firstpos [C] u32       minimum batch position that touched the slot in the running batch (0xffffffff idle)
// (Three parallel arrays on purpose: `vals` and `firstpos` of a 2^24-slot table are 64 MB each and live in the 126 MB L2
// across the five phases of a batch.  One 32-byte struct per slot was measured 12-20 % slower -- BFS depth 7: 2.75 vs
// 2.45 ms, 2^22 inserts: 0.88 vs 0.73 ms -- because every phase then misses to DRAM.)
// Keys: 20x24 -> 20 cubies x 5 bit = 100 bit (lo = cubies 0-11, hi = cubies 12-19).
//       6x8x6 -> per face the 8 sticker colours as a base-6 number (< 6^8 < 2^21), faces 0-2 in lo, 3-5 in hi;
//       injective on valid (one-hot) states.
//
// Batch insert = 5 small launches, no host synchronisation:
//   probe  : every item finds or claims its slot, atomicMin(firstpos[slot], position)
//   flag   : seen = vals[slot] != 0; first = firstpos[slot] == position; per-block count of new = first & !seen
//   scan   : exclusive scan of the block counts (one block)
//   assign : new items get index count + rank + 1 in batch order (vals[slot] = index) and are compacted
//   finish : index[i] = vals[slot_i]; firstpos reset; count += number of new states
#pragma once
#include "rb_common.cuh"

namespace rbf {

constexpr int kThreads = 256;
constexpr unsigned long long kEmpty = ~0ull;

struct Key { unsigned long long lo, hi; };

struct Table {
	ulonglong2* keys;
	int32_t* vals;
	uint32_t* firstpos;
	uint64_t mask;
};

__host__ __device__ inline Table table_of(void* base, int64_t capacity) {
	Table t;
	t.keys = reinterpret_cast<ulonglong2*>(base);
	t.vals = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(base) + capacity * 16);
	t.firstpos = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(base) + capacity * 20);
	t.mask = (uint64_t)capacity - 1;
	return t;
}

__device__ __forceinline__ uint64_t hash_key(Key k) {
	uint64_t h = k.lo * 0x9E3779B97F4A7C15ull ^ (k.hi + 0x7F4A7C15ull) * 0xC2B2AE3D27D4EB4Full;
	h ^= h >> 32;
	h *= 0xD6E8FEB86659FD93ull;
	h ^= h >> 29;
	return h;
}

// 4 state bytes (values < 32) -> 20 bits
__device__ __forceinline__ uint32_t pack4x5(uint32_t w) {
	w &= 0x1f1f1f1fu;
	w = (w & 0x001f001fu) | ((w & 0x1f001f00u) >> 3);          // two 10-bit fields at bits 0 and 16
	return (w & 0x3ffu) | ((w >> 6) & 0xffc00u);
}

__device__ __forceinline__ Key pack2024(const uint32_t (&w)[5]) {
	Key k;
	k.lo = (uint64_t)pack4x5(w[0]) | ((uint64_t)pack4x5(w[1]) << 20) | ((uint64_t)pack4x5(w[2]) << 40);
	k.hi = (uint64_t)pack4x5(w[3]) | ((uint64_t)pack4x5(w[4]) << 20);
	return k;
}

__device__ __forceinline__ ulonglong2 cas128(ulonglong2* addr, ulonglong2 cmp, ulonglong2 val) {
	ulonglong2 old;
	asm volatile(
		"{\n\t"
		".reg .b128 c, v, o;\n\t"
		"mov.b128 c, {%2, %3};\n\t"
		"mov.b128 v, {%4, %5};\n\t"
		"atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
		"mov.b128 {%0, %1}, o;\n\t"
		"}"
		: "=l"(old.x), "=l"(old.y)
		: "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(addr)
		: "memory");
	return old;
}

__device__ __forceinline__ ulonglong2 ld128_volatile(const ulonglong2* addr) {
	ulonglong2 v;
	asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(addr) : "memory");
	return v;
}

// Find the slot holding `k`, claiming an empty one if absent.  Returns -1 when the table is full.
__device__ __forceinline__ int64_t find_or_claim(const Table& t, Key k) {
	uint64_t s = hash_key(k) & t.mask;
	for (uint64_t probes = 0; probes <= t.mask; ++probes, s = (s + 1) & t.mask) {
		ulonglong2 cur = ld128_volatile(t.keys + s);
		if (cur.x == kEmpty && cur.y == kEmpty) {
			cur = cas128(t.keys + s, make_ulonglong2(kEmpty, kEmpty), make_ulonglong2(k.lo, k.hi));
			if (cur.x == kEmpty && cur.y == kEmpty) return (int64_t)s;
		}
		if (cur.x == k.lo && cur.y == k.hi) return (int64_t)s;
	}
	return -1;
}

__device__ __forceinline__ int64_t find_only(const Table& t, Key k) {
	uint64_t s = hash_key(k) & t.mask;
	for (uint64_t probes = 0; probes <= t.mask; ++probes, s = (s + 1) & t.mask) {
		const ulonglong2 cur = t.keys[s];
		if (cur.x == k.lo && cur.y == k.hi) return (int64_t)s;
		if (cur.x == kEmpty && cur.y == kEmpty) return -1;
	}
	return -1;
}

__global__ void __launch_bounds__(kThreads) k_clear(void* base, int64_t capacity) {
	const Table t = table_of(base, capacity);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += (int64_t)gridDim.x * blockDim.x) {
		t.keys[i] = make_ulonglong2(kEmpty, kEmpty);
		t.vals[i] = 0;
		t.firstpos[i] = 0xffffffffu;
	}
}

// Move every (key, index) pair of `src` into the (cleared, larger) table `dst`.
__global__ void __launch_bounds__(kThreads) k_rehash(void* src_base, int64_t src_cap, void* dst_base, int64_t dst_cap) {
	const Table s = table_of(src_base, src_cap), d = table_of(dst_base, dst_cap);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src_cap; i += (int64_t)gridDim.x * blockDim.x) {
		const ulonglong2 k = s.keys[i];
		if (k.x == kEmpty && k.y == kEmpty) continue;
		const int64_t slot = find_or_claim(d, Key{k.x, k.y});
		if (slot >= 0) d.vals[slot] = s.vals[i];
	}
}

// ---- state providers: item i -> 20x24 state words --------------------------------------------------
struct FromArray2024 {           // item i = states[i]
	const int8_t* states;
	__device__ __forceinline__ void load(int64_t i, const uint8_t*, uint32_t (&w)[5]) const {
		const uint8_t* p = reinterpret_cast<const uint8_t*>(states) + i * 20;
		if ((reinterpret_cast<uintptr_t>(states) & 3u) == 0) {
#pragma unroll
			for (int k = 0; k < 5; ++k) w[k] = reinterpret_cast<const uint32_t*>(p)[k];
		} else {
#pragma unroll
			for (int k = 0; k < 5; ++k) w[k] = p[4 * k] | (p[4 * k + 1] << 8) | (p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
		}
	}
};
struct FromParent2024 {          // item i = action (i % 12) applied to frontier[i / 12]
	const int8_t* frontier;
	__device__ __forceinline__ void load(int64_t i, const uint8_t* s_lut, uint32_t (&w)[5]) const {
		FromArray2024{frontier}.load(i / 12, s_lut, w);
		rb_move2024(s_lut, (uint32_t)(i % 12), w);
	}
};

// scratch layout for a batch of n items
struct Scratch {
	int32_t* slot;        // [n]   slot of every item (-1 = table full)
	int32_t* block_new;   // [nb + 1] per-block count of new items, then exclusive offsets; [nb] = total
	ulonglong2* keys;     // [n]   6x8x6 only: packed keys
	int64_t nb;
};
__host__ __device__ inline int64_t n_blocks(int64_t n) { return (n + kThreads - 1) / kThreads; }
__host__ __device__ inline Scratch scratch_of(void* base, int64_t n) {
	Scratch s;
	s.nb = n_blocks(n);
	s.slot = reinterpret_cast<int32_t*>(base);
	s.block_new = s.slot + ((n + 3) / 4) * 4;
	s.keys = reinterpret_cast<ulonglong2*>(s.block_new + ((s.nb + 1 + 3) / 4) * 4);
	return s;
}
inline int64_t scratch_bytes(int64_t n, bool with_keys) {
	return ((n + 3) / 4) * 16 + ((n_blocks(n) + 1 + 3) / 4) * 16 + (with_keys ? n * 16 : 0) + 16;
}

template <class Provider>
__global__ void __launch_bounds__(kThreads)
k_probe2024(Provider prov, void* base, int64_t capacity, int64_t n, int32_t* __restrict__ slot) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	rb_stage_lut2024(s_lut);
	__syncthreads();
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	uint32_t w[5];
	prov.load(i, s_lut, w);
	const int64_t s = find_or_claim(t, pack2024(w));
	slot[i] = (int32_t)s;
	if (s >= 0) atomicMin(t.firstpos + s, (uint32_t)i);
}

__global__ void __launch_bounds__(kThreads)
k_probe_keys(const ulonglong2* __restrict__ keys, void* base, int64_t capacity, int64_t n, int32_t* __restrict__ slot) {
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	const ulonglong2 k = keys[i];
	const int64_t s = find_or_claim(t, Key{k.x, k.y});
	slot[i] = (int32_t)s;
	if (s >= 0) atomicMin(t.firstpos + s, (uint32_t)i);
}

// 6x8x6 packing: warp per state; lane = one of 48 stickers (two rounds), colour = index of the set byte.
__global__ void __launch_bounds__(kThreads)
k_pack686(const int8_t* __restrict__ states, int64_t n, ulonglong2* __restrict__ keys) {
	const int lane = threadIdx.x & 31;
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	if (warp >= n) return;
	const uint8_t* p = reinterpret_cast<const uint8_t*>(states) + warp * 288;
	// lane l < 24 handles stickers 2l and 2l+1 (12 bytes = 3 words)
	uint32_t part = 0;            // c0 + 6*c1 for the lane's sticker pair
	if (lane < 24) {
		const uint32_t* q = reinterpret_cast<const uint32_t*>(p + lane * 12);
		const uint32_t a = q[0], b = q[1], c = q[2];
		uint32_t c0 = 0, c1 = 0;
		// bytes 0-5 = sticker 2l, bytes 6-11 = sticker 2l+1
		const uint32_t by[12] = {a & 0xff, (a >> 8) & 0xff, (a >> 16) & 0xff, a >> 24, b & 0xff, (b >> 8) & 0xff,
		                         (b >> 16) & 0xff, b >> 24, c & 0xff, (c >> 8) & 0xff, (c >> 16) & 0xff, c >> 24};
#pragma unroll
		for (int k = 0; k < 6; ++k) { c0 += by[k] ? k : 0; c1 += by[6 + k] ? k : 0; }
		part = (c0 % 6u) + 6u * (c1 % 6u);
	}
	// face f = lanes 4f..4f+3: value = sum part_j * 36^j
	const uint32_t p1 = __shfl_down_sync(0xffffffffu, part, 1);
	const uint32_t p2 = __shfl_down_sync(0xffffffffu, part, 2);
	const uint32_t p3 = __shfl_down_sync(0xffffffffu, part, 3);
	const uint32_t face = part + 36u * p1 + 1296u * p2 + 46656u * p3;      // < 6^8, valid on lanes 0,4,...,20
	const uint64_t f0 = __shfl_sync(0xffffffffu, face, 0), f1 = __shfl_sync(0xffffffffu, face, 4);
	const uint64_t f2 = __shfl_sync(0xffffffffu, face, 8), f3 = __shfl_sync(0xffffffffu, face, 12);
	const uint64_t f4 = __shfl_sync(0xffffffffu, face, 16), f5 = __shfl_sync(0xffffffffu, face, 20);
	if (lane == 0) keys[warp] = make_ulonglong2(f0 | (f1 << 21) | (f2 << 42), f3 | (f4 << 21) | (f5 << 42));
}

// flags + per-block count of new items
__global__ void __launch_bounds__(kThreads)
k_flag(void* base, int64_t capacity, int64_t n, const int32_t* __restrict__ slot, uint8_t* __restrict__ seen,
       uint8_t* __restrict__ first, int32_t* __restrict__ block_new) {
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	bool is_new = false;
	if (i < n) {
		const int32_t s = slot[i];
		bool sn = false, fs = false;
		if (s >= 0) {
			sn = t.vals[s] != 0;
			fs = t.firstpos[s] == (uint32_t)i;
		}
		if (seen) seen[i] = sn;
		if (first) first[i] = fs;
		is_new = fs && !sn;
	}
	const int c = __syncthreads_count(is_new);
	if (threadIdx.x == 0) block_new[blockIdx.x] = c;
}

// exclusive scan of block counts in place; block_new[nb] = total.  One block.
__global__ void __launch_bounds__(1024) k_scan(int32_t* __restrict__ block_new, int64_t nb) {
	__shared__ int32_t s_warp[32];
	__shared__ int32_t s_carry;
	if (threadIdx.x == 0) s_carry = 0;
	__syncthreads();
	for (int64_t base = 0; base < nb; base += 1024) {
		const int64_t i = base + threadIdx.x;
		const int32_t x = i < nb ? block_new[i] : 0;
		int32_t v = x;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const int32_t y = __shfl_up_sync(0xffffffffu, v, o);
			if ((threadIdx.x & 31) >= o) v += y;
		}
		if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = v;
		__syncthreads();
		if (threadIdx.x < 32) {
			int32_t wv = s_warp[threadIdx.x];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const int32_t y = __shfl_up_sync(0xffffffffu, wv, o);
				if (threadIdx.x >= o) wv += y;
			}
			s_warp[threadIdx.x] = wv;
		}
		__syncthreads();
		const int32_t carry = s_carry;
		const int32_t incl = v + (threadIdx.x >= 32 ? s_warp[(threadIdx.x >> 5) - 1] : 0);
		if (i < nb) block_new[i] = carry + incl - x;
		__syncthreads();
		if (threadIdx.x == 1023) s_carry = carry + incl;
		__syncthreads();
	}
	if (threadIdx.x == 0) block_new[nb] = s_carry;
}

// Rank of a new item inside its block (batch order), via ballot + warp prefix.
__device__ __forceinline__ int block_rank(bool flag, int32_t* s_warp) {
	const unsigned m = __ballot_sync(0xffffffffu, flag);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (lane == 0) s_warp[wid] = __popc(m);
	__syncthreads();
	int off = 0;
	for (int k = 0; k < wid; ++k) off += s_warp[k];
	return off + __popc(m & ((1u << lane) - 1u));
}

// assign indices to new items (batch order) and compact them.
//   2024 provider given: writes next_frontier (state), parent, action, solved for each new item at its rank.
template <class Provider, bool kFrontier>
__global__ void __launch_bounds__(kThreads)
k_assign2024(Provider prov, void* base, int64_t capacity, int64_t n, const int32_t* __restrict__ slot,
             const int32_t* __restrict__ block_new, const int32_t* __restrict__ count, int8_t* __restrict__ next_frontier,
             int32_t* __restrict__ parent, uint8_t* __restrict__ action, uint8_t* __restrict__ solved) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	__shared__ int32_t s_warp[kThreads / 32];
	if (kFrontier) rb_stage_lut2024(s_lut);
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	int32_t s = -1;
	bool is_new = false;
	if (i < n) {
		s = slot[i];
		if (s >= 0) is_new = t.vals[s] == 0 && t.firstpos[s] == (uint32_t)i;
	}
	const int r = block_rank(is_new, s_warp);      // also orders the LUT staging before use
	if (!is_new) return;
	const int64_t k = (int64_t)block_new[blockIdx.x] + r;
	t.vals[s] = *count + (int32_t)k + 1;
	if (kFrontier) {
		uint32_t w[5];
		prov.load(i, s_lut, w);
		if (next_frontier) {
			uint32_t* dst = reinterpret_cast<uint32_t*>(next_frontier + k * 20);     // 20 B records: 4-byte aligned
#pragma unroll
			for (int q = 0; q < 5; ++q) dst[q] = w[q];
		}
		if (parent) parent[k] = (int32_t)(i / 12);
		if (action) action[k] = (uint8_t)(i % 12);
		if (solved) {
			const uint32_t* sv = reinterpret_cast<const uint32_t*>(g_solved2024);
			solved[k] = (w[0] == sv[0]) & (w[1] == sv[1]) & (w[2] == sv[2]) & (w[3] == sv[3]) & (w[4] == sv[4]);
		}
	}
}

// generic assign on precomputed slots (no state provider): only vals + optional compaction of item ids
__global__ void __launch_bounds__(kThreads)
k_assign_ids(void* base, int64_t capacity, int64_t n, const int32_t* __restrict__ slot, const int32_t* __restrict__ block_new,
             const int32_t* __restrict__ count, int32_t* __restrict__ new_items) {
	__shared__ int32_t s_warp[kThreads / 32];
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	int32_t s = -1;
	bool is_new = false;
	if (i < n) {
		s = slot[i];
		if (s >= 0) is_new = t.vals[s] == 0 && t.firstpos[s] == (uint32_t)i;
	}
	const int r = block_rank(is_new, s_warp);
	if (!is_new) return;
	const int64_t k = (int64_t)block_new[blockIdx.x] + r;
	t.vals[s] = *count + (int32_t)k + 1;
	if (new_items) new_items[k] = (int32_t)i;
}

// index[i] = vals[slot_i]; firstpos reset; count += total (after every block has read it: done by a second tiny launch)
__global__ void __launch_bounds__(kThreads)
k_finish(void* base, int64_t capacity, int64_t n, const int32_t* __restrict__ slot, int32_t* __restrict__ index) {
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	const int32_t s = slot[i];
	if (index) index[i] = s >= 0 ? t.vals[s] : -1;
	if (s >= 0) t.firstpos[s] = 0xffffffffu;
}

__global__ void k_bump(int32_t* __restrict__ count, const int32_t* __restrict__ block_new, int64_t nb, int32_t* __restrict__ n_new) {
	const int32_t total = block_new[nb];
	if (n_new) *n_new = total;
	*count += total;
}

__global__ void __launch_bounds__(kThreads)
k_lookup2024(const int8_t* __restrict__ states, const void* base, int64_t capacity, int64_t n, int32_t* __restrict__ index) {
	const Table t = table_of(const_cast<void*>(base), capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	uint32_t w[5];
	FromArray2024{states}.load(i, nullptr, w);
	const int64_t s = find_only(t, pack2024(w));
	index[i] = s >= 0 ? t.vals[s] : 0;
}

__global__ void __launch_bounds__(kThreads)
k_lookup_keys(const ulonglong2* __restrict__ keys, const void* base, int64_t capacity, int64_t n, int32_t* __restrict__ index) {
	const Table t = table_of(const_cast<void*>(base), capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	const ulonglong2 k = keys[i];
	const int64_t s = find_only(t, Key{k.x, k.y});
	index[i] = s >= 0 ? t.vals[s] : 0;
}

// 6x8x6 frontier compaction: gather the new children (ids in new_items) into next_frontier, with parent/action/solved.
__global__ void __launch_bounds__(kThreads)
k_gather686(const int8_t* __restrict__ children, const int32_t* __restrict__ new_items, const int32_t* __restrict__ block_new,
            int64_t nb, int8_t* __restrict__ next_frontier, int32_t* __restrict__ parent, uint8_t* __restrict__ action,
            uint8_t* __restrict__ solved) {
	const int lane = threadIdx.x & 31;
	const int64_t n_new = block_new[nb];
	const int64_t warp = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
	const int64_t n_warps = (int64_t)gridDim.x * (kThreads / 32);
	for (int64_t k = warp; k < n_new; k += n_warps) {
		const int32_t i = new_items[k];
		const uint4* src = reinterpret_cast<const uint4*>(children + (int64_t)i * 288);
		uint4 v = make_uint4(0, 0, 0, 0);
		if (lane < 18) {
			v = src[lane];
			if (next_frontier) reinterpret_cast<uint4*>(next_frontier + k * 288)[lane] = v;
		}
		if (solved) {
			bool ok = true;
			if (lane < 18) {
				const uint4 sv = reinterpret_cast<const uint4*>(g_solved686)[lane];
				ok = v.x == sv.x && v.y == sv.y && v.z == sv.z && v.w == sv.w;
			}
			ok = __all_sync(0xffffffffu, ok);
			if (lane == 0) solved[k] = ok;
		}
		if (lane == 0) {
			if (parent) parent[k] = i / 12;
			if (action) action[k] = (uint8_t)(i % 12);
		}
	}
}

}  // namespace rbf
