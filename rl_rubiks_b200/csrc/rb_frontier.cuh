// Search-frontier kernels: open-addressing hash set on the packed state with the reference's batch-order
// index semantics, 12-neighbour frontier expansion with dedup and in-order compaction.
//
// Replaces the Python dict keyed on `state.tostring()` of BFS / AStar / MCTS
// (reference: librubiks/solving/agents.py:103-121, 286-306, 517-526, 605-609).
//
// Table layout in caller-owned HBM: capacity C slots (C a power of two <= 2^30), one 32-byte slot per state
//   key      2 x u64   packed state; all-ones = empty.  Claimed with one 128-bit atom.cas.
//   val      i32       1-based index of the state (insertion order); 0 = claimed in the running batch, not numbered yet
//   firstpos u32       minimum batch position that touched the slot in the running batch (0xffffffff idle)
//   (8 bytes spare)
// A slot is one 32-byte sector, so every phase of a batch costs ONE random DRAM burst per item.  (Round 1 kept three parallel
// arrays -- keys / vals / firstpos -- and five phases; ncu showed 140 B of DRAM reads per item and phase, two 64-byte bursts,
// 5.3 GB per depth-7 BFS closure against 0.7 GB algorithmic: profiles/r1k_frontier_bfs7_launches.txt.)
// Keys: 20x24 -> 20 cubies x 5 bit = 100 bit (lo = cubies 0-11, hi = cubies 12-19).
//       6x8x6 -> per face the 8 sticker colours as a base-6 number (< 6^8 < 2^21), faces 0-2 in lo, 3-5 in hi;
//       injective on valid (one-hot) states.
//
// Batch insert = 2 launches (3 when the caller wants every item's index), no host synchronisation:
//   probe   : every item finds or claims its slot; seen = val != 0 (states numbered by earlier batches); items not seen
//             (and seen ones too when the caller asks for the `first` flags) do old = atomicMin(firstpos, position) and, when
//             the slot was already in the race, mark the LOSER max(old, position) in a byte-per-item scratch array: at the end
//             of the pass exactly the minimum position of every slot is unmarked -- the first occurrence is known without
//             reading the table again.  The slot number and the seen bit go to a scratch word per item; seen items already
//             know their index
//   resolve : first = not marked; new = first & !seen; ONE pass numbers the new items in batch order -- a decoupled
//             look-back scan over the blocks' counts (each block publishes its count, then its inclusive prefix; blocks take
//             their number from a ticket so that every predecessor is resident or done) -- stores val = count + rank + 1 and
//             re-arms firstpos (one 8-byte store into the slot, no table read in this pass), compacts the new items (state,
//             parent, action, solved) and, in the last block, count += total
//   index   : (optional) items that were not seen read their state's index back
// A full table (a probe that wraps around) sets an error word: count (and n_new) become -1 and stay so -- the caller sees
// RB_ERR_CAPACITY semantics without a host sync inside the call (the Python mirror raises on the next read of the size).
#pragma once
#include "rb_common.cuh"

namespace rbf {

constexpr int kThreads = 256;
constexpr unsigned long long kEmpty = ~0ull;
constexpr int64_t kMaxCapacity = int64_t(1) << 30;          // slot numbers travel in 31 bits (+ the seen bit)
constexpr uint32_t kNoSlot = 0xffffffffu;                   // scratch word of an item dropped by a full table

struct Key { unsigned long long lo, hi; };

struct __align__(32) Slot {
	unsigned long long lo, hi;
	int32_t val;
	uint32_t firstpos;
	unsigned long long spare;
};

struct Table {
	Slot* slots;
	uint64_t mask;
	__device__ __forceinline__ ulonglong2* key(int64_t s) const { return reinterpret_cast<ulonglong2*>(slots + s); }
	__device__ __forceinline__ int32_t& val(int64_t s) const { return slots[s].val; }
	__device__ __forceinline__ uint32_t& firstpos(int64_t s) const { return slots[s].firstpos; }
};

__host__ __device__ inline Table table_of(void* base, int64_t capacity) {
	Table t;
	t.slots = reinterpret_cast<Slot*>(base);
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
__device__ __forceinline__ int32_t ld32_volatile(const int32_t* addr) {
	int32_t v;
	asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
	return v;
}

// Find the slot holding `k`, claiming an empty one if absent.  Returns -1 when the table is full; *claimed = this call put the
// key there (the slot is fresh: val 0, firstpos idle).
__device__ __forceinline__ int64_t find_or_claim(const Table& t, Key k, bool* claimed = nullptr) {
	uint64_t s = hash_key(k) & t.mask;
	if (claimed) *claimed = false;
	for (uint64_t probes = 0; probes <= t.mask; ++probes, s = (s + 1) & t.mask) {
		ulonglong2 cur = ld128_volatile(t.key(s));
		if (cur.x == kEmpty && cur.y == kEmpty) {
			cur = cas128(t.key(s), make_ulonglong2(kEmpty, kEmpty), make_ulonglong2(k.lo, k.hi));
			if (cur.x == kEmpty && cur.y == kEmpty) {
				if (claimed) *claimed = true;
				return (int64_t)s;
			}
		}
		if (cur.x == k.lo && cur.y == k.hi) return (int64_t)s;
	}
	return -1;
}

__device__ __forceinline__ int64_t find_only(const Table& t, Key k) {
	uint64_t s = hash_key(k) & t.mask;
	for (uint64_t probes = 0; probes <= t.mask; ++probes, s = (s + 1) & t.mask) {
		const ulonglong2 cur = *t.key(s);
		if (cur.x == k.lo && cur.y == k.hi) return (int64_t)s;
		if (cur.x == kEmpty && cur.y == kEmpty) return -1;
	}
	return -1;
}

__global__ void __launch_bounds__(kThreads) k_clear(void* base, int64_t capacity) {
	uint4* p = reinterpret_cast<uint4*>(base);                    // slot = {key all-ones} {val 0, firstpos idle, spare 0}
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * capacity; i += (int64_t)gridDim.x * blockDim.x)
		p[i] = (i & 1) ? make_uint4(0u, 0xffffffffu, 0u, 0u) : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
}

// Move every (key, index) pair of `src` into the (cleared, larger) table `dst`.
__global__ void __launch_bounds__(kThreads) k_rehash(void* src_base, int64_t src_cap, void* dst_base, int64_t dst_cap) {
	const Table s = table_of(src_base, src_cap), d = table_of(dst_base, dst_cap);
	for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < src_cap; i += (int64_t)gridDim.x * blockDim.x) {
		const ulonglong2 k = *s.key(i);
		if (k.x == kEmpty && k.y == kEmpty) continue;
		const int64_t slot = find_or_claim(d, Key{k.x, k.y});
		if (slot >= 0) d.val(slot) = s.val(i);
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
		const uint32_t u = (uint32_t)i, par = u / 12u;           // item numbers fit 31 bits (checked by the entry points): no 64-bit division
		FromArray2024{frontier}.load(par, s_lut, w);
		rb_move2024(s_lut, u - 12u * par, w);
	}
};

// scratch layout for a batch of n items
struct Scratch {
	uint32_t* word;               // [n]  slot | seen << 31 of every item; kNoSlot = dropped by a full table
	uint8_t* lost;                // [n]  1 = another item of this batch with a smaller position has the same state
	unsigned long long* status;   // [nb] look-back state of the resolve pass: flag << 32 | count (flag 1 = block count, 2 = inclusive prefix)
	uint32_t* ctl;                // [4]  ticket, error word, 2 spare (lost, status and ctl are zeroed by ONE memset per batch)
	ulonglong2* keys;             // [n]  6x8x6 only: packed keys
	int64_t nb;
};
__host__ __device__ inline int64_t n_blocks(int64_t n) { return (n + kThreads - 1) / kThreads; }
__host__ __device__ inline int64_t up16(int64_t x) { return (x + 15) / 16 * 16; }
__host__ __device__ inline Scratch scratch_of(void* base, int64_t n) {
	Scratch s;
	s.nb = n_blocks(n);
	uint8_t* p = reinterpret_cast<uint8_t*>(base);
	s.word = reinterpret_cast<uint32_t*>(p); p += up16(n * 4);
	s.lost = p; p += up16(n);
	s.status = reinterpret_cast<unsigned long long*>(p); p += up16(s.nb * 8);
	s.ctl = reinterpret_cast<uint32_t*>(p); p += 16;
	s.keys = reinterpret_cast<ulonglong2*>(p);
	return s;
}
inline int64_t scratch_bytes(int64_t n, bool with_keys) { return up16(n * 4) + up16(n) + up16(n_blocks(n) * 8) + 16 + (with_keys ? n * 16 : 0) + 16; }
inline int64_t control_bytes(int64_t n) { return up16(n) + up16(n_blocks(n) * 8) + 16; }       // lost + status + ctl: zeroed before a batch (from .lost)

// ---- probe ------------------------------------------------------------------------------------------------------------------
// One item: find or claim the slot, read whether an earlier batch numbered the state, join the race for "first occurrence in
// this batch" and leave slot | seen << 31 in the scratch word.
__device__ __forceinline__ void probe_item(const Table& t, Key k, int64_t i, bool need_first, uint32_t* __restrict__ word, uint8_t* lost,
                                           uint32_t* __restrict__ ctl, uint8_t* __restrict__ seen, int32_t* __restrict__ index) {
	bool claimed;
	const int64_t s = find_or_claim(t, k, &claimed);
	if (s < 0) {
		word[i] = kNoSlot;
		atomicOr(ctl + 1, 1u);
		if (seen) seen[i] = 0;
		if (index) index[i] = -1;
		return;
	}
	// numbered by an earlier batch?  (vals of this batch are written by resolve; a slot this item claimed is fresh: no read)
	const int32_t v = claimed ? 0 : ld32_volatile(&t.val(s));
	const bool sn = v != 0;
	if (!sn || need_first) {
		const uint32_t old = atomicMin(&t.firstpos(s), (uint32_t)i);      // the sector was just read / claimed: an L2 hit
		if (old != 0xffffffffu) lost[old > (uint32_t)i ? old : (uint32_t)i] = 1;
	}
	word[i] = (uint32_t)s | (sn ? 0x80000000u : 0u);
	if (seen) seen[i] = sn;
	if (index && sn) index[i] = v;
}

template <class Provider>
__global__ void __launch_bounds__(kThreads)
k_probe2024(Provider prov, void* base, int64_t capacity, int64_t n, int need_first, uint32_t* __restrict__ word, uint8_t* lost,
            uint32_t* __restrict__ ctl, uint8_t* __restrict__ seen, int32_t* __restrict__ index, const int32_t* __restrict__ n_units_dev, int per_unit) {
	// n_units_dev != nullptr: the batch really holds per_unit * *n_units_dev items (<= n, the size the grid was launched for): lets a
	// caller chain batches whose size is only known on the device (the layers of a BFS) without a host round trip
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	rb_stage_lut2024(s_lut);
	__syncthreads();
	if (n_units_dev) n = min(n, (int64_t)per_unit * (int64_t)*n_units_dev);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	uint32_t w[5];
	prov.load(i, s_lut, w);
	probe_item(table_of(base, capacity), pack2024(w), i, need_first != 0, word, lost, ctl, seen, index);
}

__global__ void __launch_bounds__(kThreads)
k_probe_keys(const ulonglong2* __restrict__ keys, void* base, int64_t capacity, int64_t n, int need_first, uint32_t* __restrict__ word,
             uint8_t* lost, uint32_t* __restrict__ ctl, uint8_t* __restrict__ seen, int32_t* __restrict__ index) {
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	const ulonglong2 k = keys[i];
	probe_item(table_of(base, capacity), Key{k.x, k.y}, i, need_first != 0, word, lost, ctl, seen, index);
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

// ---- resolve ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long* p) {
	unsigned long long v;
	asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_status(unsigned long long* p, unsigned long long v) {
	asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of this block's count over all earlier blocks (decoupled look-back, Merrill & Garland): warp 0 publishes
// the block's count, walks back over the predecessors' status words 32 at a time until it meets an inclusive prefix, then
// publishes its own inclusive prefix.  `b` is the block's ticket, so every predecessor is running or done.
__device__ __forceinline__ int32_t lookback(unsigned long long* status, int32_t b, int32_t agg) {
	const int lane = threadIdx.x & 31;
	if (lane == 0) st_status(status + b, ((unsigned long long)(b == 0 ? 2 : 1) << 32) | (uint32_t)agg);
	int32_t excl = 0;
	for (int32_t j = b - 1; j >= 0;) {
		const int32_t idx = j - lane;
		const unsigned long long v = idx >= 0 ? ld_status(status + idx) : (2ull << 32);
		const uint32_t flag = (uint32_t)(v >> 32);
		const unsigned ready = __ballot_sync(0xffffffffu, flag != 0), incl = __ballot_sync(0xffffffffu, flag == 2);
		const int stop = incl ? __ffs(incl) - 1 : 31;                       // nearest predecessor that already has its inclusive prefix
		const unsigned need = stop == 31 ? 0xffffffffu : ((2u << stop) - 1u);
		if ((ready & need) != need) continue;                               // some predecessor in the window has not published yet
		int32_t c = lane <= stop ? (int32_t)(uint32_t)v : 0;
#pragma unroll
		for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
		excl += c;
		if (incl) break;
		j -= 32;
	}
	if (lane == 0 && b > 0) st_status(status + b, (2ull << 32) | (uint32_t)(excl + agg));
	return excl;
}

// Rank of a new item inside its block (batch order), via ballot + warp prefix; *total = the block's count.
__device__ __forceinline__ int block_rank(bool flag, int32_t* s_warp, int32_t* total = nullptr) {
	const unsigned m = __ballot_sync(0xffffffffu, flag);
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (lane == 0) s_warp[wid] = __popc(m);
	__syncthreads();
	int off = 0, all = 0;
#pragma unroll
	for (int k = 0; k < kThreads / 32; ++k) {
		const int c = s_warp[k];
		off += k < wid ? c : 0;
		all += c;
	}
	if (total) *total = all;
	return off + __popc(m & ((1u << lane) - 1u));
}

// first / new flags, numbering of the new items in batch order, compaction -- one pass.
//   kFrontier: the 2024 provider regenerates the item's state; new item k gets next_frontier / parent / action / solved row k.
//   otherwise: new_items[k] = item number (6x8x6 frontier: gathered afterwards) or nothing (plain insert).
template <class Provider, bool kFrontier>
__global__ void __launch_bounds__(kThreads)
k_resolve(Provider prov, void* base, int64_t capacity, int64_t n, int need_first, const uint32_t* __restrict__ word,
          const uint8_t* __restrict__ lost, unsigned long long* __restrict__ status, uint32_t* __restrict__ ctl, int32_t* __restrict__ count, int32_t* __restrict__ n_new,
          uint8_t* __restrict__ first, int32_t* __restrict__ index, int32_t* __restrict__ new_items, int8_t* __restrict__ next_frontier,
          int32_t* __restrict__ parent, uint8_t* __restrict__ action, uint8_t* __restrict__ solved, const int32_t* __restrict__ n_units_dev, int per_unit) {
	__shared__ __align__(16) uint8_t s_lut[RB_LUT_BYTES];
	__shared__ int32_t s_warp[kThreads / 32];
	__shared__ int32_t s_bid, s_excl, s_count;
	if (kFrontier) rb_stage_lut2024(s_lut);
	if (threadIdx.x == 0) s_bid = (int32_t)atomicAdd(ctl, 1u);
	__syncthreads();
	const int32_t b = s_bid;
	const Table t = table_of(base, capacity);
	const int64_t i = (int64_t)b * kThreads + threadIdx.x;
	const int64_t n_launched = n;                                             // the grid covers this many items; fewer may exist (n_units_dev)
	if (n_units_dev) n = min(n, (int64_t)per_unit * (int64_t)*n_units_dev);
	uint32_t wd = kNoSlot;
	bool fs = false, is_new = false;
	if (i < n) {
		wd = word[i];
		if (wd != kNoSlot) {
			const bool sn = (wd >> 31) != 0;
			fs = (!sn || need_first) && !lost[i];                                  // no smaller position met this item's slot in the probe pass
			is_new = fs && !sn;
		}
		if (first) first[i] = fs;
	}
	int32_t agg;
	const int r = block_rank(is_new, s_warp, &agg);
	if (threadIdx.x < 32) {
		const int32_t c0 = *reinterpret_cast<volatile int32_t*>(count);          // read before this block publishes: the last block
		const int32_t excl = lookback(status, b, agg);                            // rewrites count only after every block has read it
		if (threadIdx.x == 0) {
			s_excl = excl; s_count = c0;
			if ((int64_t)b == (n_launched + kThreads - 1) / kThreads - 1) {
				const bool bad = c0 < 0 || *reinterpret_cast<volatile uint32_t*>(ctl + 1) != 0;
				*count = bad ? -1 : c0 + excl + agg;
				if (n_new) *n_new = bad ? -1 : excl + agg;
			}
		}
	}
	__syncthreads();
	if (fs && !is_new) t.firstpos(wd & 0x7fffffffu) = 0xffffffffu;                // a seen slot that was raced for: re-armed by its winner
	if (!is_new) return;
	const int64_t k = (int64_t)s_excl + r;
	const int32_t idx = s_count + (int32_t)k + 1;
	// val and the re-armed firstpos in one 8-byte store: the only table access of this pass
	*reinterpret_cast<uint2*>(&t.slots[wd & 0x7fffffffu].val) = make_uint2((uint32_t)idx, 0xffffffffu);
	if (index) index[i] = idx;
	if (new_items) new_items[k] = (int32_t)i;
	if (kFrontier) {
		const uint32_t u = (uint32_t)i, par = u / 12u;
		uint32_t w[5] = {0u, 0u, 0u, 0u, 0u};
		if (next_frontier || solved) prov.load(i, s_lut, w);                    // the state itself is regenerated only when somebody wants it
		if (next_frontier) {
			uint32_t* dst = reinterpret_cast<uint32_t*>(next_frontier + k * 20);     // 20 B records: 4-byte aligned
#pragma unroll
			for (int q = 0; q < 5; ++q) dst[q] = w[q];
		}
		if (parent) parent[k] = (int32_t)par;
		if (action) action[k] = (uint8_t)(u - 12u * par);
		if (solved) {
			const uint32_t* sv = reinterpret_cast<const uint32_t*>(g_solved2024);
			solved[k] = (w[0] == sv[0]) & (w[1] == sv[1]) & (w[2] == sv[2]) & (w[3] == sv[3]) & (w[4] == sv[4]);
		}
	}
}

struct NoProvider {
	__device__ __forceinline__ void load(int64_t, const uint8_t*, uint32_t (&)[5]) const {}
};

// index of the items that were claimed in this batch but are not its first occurrence (the others know theirs already)
__global__ void __launch_bounds__(kThreads)
k_index_rest(const void* base, int64_t capacity, int64_t n, const uint32_t* __restrict__ word, const uint8_t* __restrict__ first,
             int32_t* __restrict__ index) {
	const Table t = table_of(const_cast<void*>(base), capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	const uint32_t wd = word[i];
	if (wd == kNoSlot || (wd >> 31) || (first && first[i])) return;
	index[i] = t.val(wd);
}

__global__ void __launch_bounds__(kThreads)
k_lookup2024(const int8_t* __restrict__ states, const void* base, int64_t capacity, int64_t n, int32_t* __restrict__ index) {
	const Table t = table_of(const_cast<void*>(base), capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	uint32_t w[5];
	FromArray2024{states}.load(i, nullptr, w);
	const int64_t s = find_only(t, pack2024(w));
	index[i] = s >= 0 ? t.val(s) : 0;
}

__global__ void __launch_bounds__(kThreads)
k_lookup_keys(const ulonglong2* __restrict__ keys, const void* base, int64_t capacity, int64_t n, int32_t* __restrict__ index) {
	const Table t = table_of(const_cast<void*>(base), capacity);
	const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
	if (i >= n) return;
	const ulonglong2 k = keys[i];
	const int64_t s = find_only(t, Key{k.x, k.y});
	index[i] = s >= 0 ? t.val(s) : 0;
}

// 6x8x6 frontier compaction: gather the new children (ids in new_items) into next_frontier, with parent/action/solved.
__global__ void __launch_bounds__(kThreads)
k_gather686(const int8_t* __restrict__ children, const int32_t* __restrict__ new_items, const int32_t* __restrict__ n_new_dev,
            int8_t* __restrict__ next_frontier, int32_t* __restrict__ parent, uint8_t* __restrict__ action,
            uint8_t* __restrict__ solved) {
	const int lane = threadIdx.x & 31;
	const int64_t n_new = *n_new_dev;
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
