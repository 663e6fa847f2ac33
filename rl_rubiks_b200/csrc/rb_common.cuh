// Shared helpers for librubiks_b200: error plumbing, device tables, vector load/store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <mutex>

#include "../../include/rubiks_b200.h"

#define RB_NUM_SMS 148          // B200: 2 dies x 74 SMs; grids are sized in multiples of this

// ---------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_rb_err[512] = "";
static std::atomic<int64_t> g_rb_launches{0};

static int rb_fail(int code, const char* fmt, const char* a = "", const char* b = "") {
	snprintf(g_rb_err, sizeof(g_rb_err), fmt, a, b);
	return code;
}

#define RB_CUDA(call)                                                                         \
	do {                                                                                      \
		cudaError_t e_ = (call);                                                              \
		if (e_ != cudaSuccess) return rb_fail(RB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
	} while (0)

#define RB_LAUNCHED(name)                                                                     \
	do {                                                                                      \
		cudaError_t e_ = cudaGetLastError();                                                  \
		if (e_ != cudaSuccess) return rb_fail(RB_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e_)); \
		g_rb_launches.fetch_add(1, std::memory_order_relaxed);                                \
	} while (0)

#define RB_REQUIRE(cond, msg)                                                                 \
	do { if (!(cond)) return rb_fail(RB_ERR_BAD_ARG, "%s (%s)", msg, #cond); } while (0)

// Grid for a grid-stride kernel: enough blocks for the work, at most `blocks_per_sm` per SM.  The caps are measured per kernel
// (profiles/r2g_grid_sweep.txt): the kernels that mostly WRITE (one-hot rows, 12-neighbour expansion, the ADI batch) are 6-18 %
// faster with one block per unit of work (kGridUncapped: the block scheduler hands out work as SMs free up, and the address streams
// of the resident warps are not locked into one fixed stride) than as a persistent grid of 8 blocks per SM; the read + write
// streaming kernels (multi_rotate) want exactly their resident block count (5 per SM at 46-47 registers: no second, partial wave).
constexpr int kGridUncapped = 1 << 20;
static inline int rb_grid(int64_t work_items, int per_block, int blocks_per_sm) {
	int64_t need = (work_items + per_block - 1) / per_block;
	int64_t cap = (int64_t)RB_NUM_SMS * blocks_per_sm;
	if (need < 1) need = 1;
	return (int)(need < cap ? need : cap);
}

// ---------------------------------------------------------------------------------------------
// device tables (generated on the host by rb_tables.cuh, uploaded once per device)
// ---------------------------------------------------------------------------------------------
// g_lut2024[(a*2 + kind)*32 + s] = new value of a cubie with value s under action a (rows padded 24 -> 32
// with the identity so that a masked out-of-range value stays in bounds).
__device__ __align__(16) uint8_t g_lut2024[12 * 2 * 32];
// The same LUT in constant memory, as words: c_lut2024[a*16 + kind*8 + k] = entries 4k..4k+3.  Kernels that apply the
// SAME action in every lane (12-neighbour expansion) read rows as constant-bank operands: no shared-memory traffic.
__constant__ uint32_t c_lut2024[12 * 16];
// g_perm686[a*48 + slot] = source sticker slot.
__device__ __align__(16) uint8_t g_perm686[12 * 48];
// Sticker tables for rendering a 6x8x6 state from a 20x24 one: [cubie][value][k] destination slot, then [cubie][k] home slot.
__device__ __align__(16) uint8_t g_stickers686[20 * 24 * 3 + 20 * 3 + 4];
// Solved states.
__device__ __align__(16) uint8_t g_solved2024[32];     // 20 used, padded to 8 words
__device__ __align__(16) uint8_t g_solved686[288];

#define RB_LUT_BYTES (12 * 2 * 32)

__device__ __forceinline__ void rb_stage_lut2024(uint8_t* s_lut) {
	// 768 B = 48 x 16 B
	for (int i = threadIdx.x; i < RB_LUT_BYTES / 16; i += blockDim.x)
		reinterpret_cast<uint4*>(s_lut)[i] = reinterpret_cast<const uint4*>(g_lut2024)[i];
}

// One move on 4 packed cubie values (one 32-bit word of the int8[20] state) through the staged LUT row.
__device__ __forceinline__ uint32_t rb_lut_word(const uint8_t* row, uint32_t w) {
	uint32_t r = row[w & 31u];
	r |= (uint32_t)row[(w >> 8) & 31u] << 8;
	r |= (uint32_t)row[(w >> 16) & 31u] << 16;
	r |= (uint32_t)row[(w >> 24) & 31u] << 24;
	return r;
}

// One move on a whole 20x24 state held as 5 words: words 0-1 are corners (kind 0), 2-4 edges (kind 1).
__device__ __forceinline__ void rb_move2024(const uint8_t* s_lut, uint32_t a, uint32_t (&w)[5]) {
	const uint8_t* rc = s_lut + a * 64u;
	const uint8_t* re = rc + 32;
	w[0] = rb_lut_word(rc, w[0]);
	w[1] = rb_lut_word(rc, w[1]);
	w[2] = rb_lut_word(re, w[2]);
	w[3] = rb_lut_word(re, w[3]);
	w[4] = rb_lut_word(re, w[4]);
}

__device__ __forceinline__ uint32_t rb_action_of(uint32_t face, uint32_t dir) {
	uint32_t a = face * 2u + (1u - (dir & 1u));
	return a < 12u ? a : 11u;
}
__device__ __forceinline__ uint32_t rb_clamp_action(uint32_t a) { return a < 12u ? a : 11u; }

// ---------------------------------------------------------------------------------------------
// streaming vector access
// ---------------------------------------------------------------------------------------------
// Store policy of the big write-once outputs: 0 = st.global.cs (streaming, evict first), 1 = plain write-back.
// Measured on B200 (profiles/r1i_configs.md vs r1g): outputs of several GB are 5-8 % faster with plain stores, outputs of a
// few hundred MB (one ADI rollout, one sequence batch) up to 17 % faster with .cs, so the launcher picks by output size
// (rb_store_policy); RB_STORE_POLICY=0|1 overrides.
#define RB_STORE_CS 0
#define RB_STORE_WB 1
__device__ __forceinline__ void rb_st_stream(float4* p, float4 v, int pol) {
	if (pol == RB_STORE_CS) __stcs(p, v);
	else *p = v;
}
__device__ __forceinline__ void rb_st_stream(uint4* p, uint4 v, int pol) {
	if (pol == RB_STORE_CS) __stcs(p, v);
	else *p = v;
}
static inline int rb_store_policy(int64_t out_bytes) {
	static const int forced = [] { const char* e = getenv("RB_STORE_POLICY"); return e ? atoi(e) : -1; }();
	if (forced == 0 || forced == 1) return forced;
	return out_bytes >= (int64_t(1) << 31) ? RB_STORE_WB : RB_STORE_CS;
}

__device__ __forceinline__ uint4 rb_ld_stream(const uint4* p) { return __ldcs(p); }

// Copy nbytes between global and shared with 16-byte vectors when the global side allows it.
__device__ __forceinline__ void rb_g2s(uint8_t* s, const uint8_t* g, int nbytes) {
	if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
		int nv = nbytes >> 4;
		for (int i = threadIdx.x; i < nv; i += blockDim.x)
			reinterpret_cast<uint4*>(s)[i] = rb_ld_stream(reinterpret_cast<const uint4*>(g) + i);
		for (int i = (nv << 4) + threadIdx.x; i < nbytes; i += blockDim.x) s[i] = g[i];
	} else if ((reinterpret_cast<uintptr_t>(g) & 3u) == 0) {
		int nv = nbytes >> 2;
		for (int i = threadIdx.x; i < nv; i += blockDim.x)
			reinterpret_cast<uint32_t*>(s)[i] = reinterpret_cast<const uint32_t*>(g)[i];
		for (int i = (nv << 2) + threadIdx.x; i < nbytes; i += blockDim.x) s[i] = g[i];
	} else {
		for (int i = threadIdx.x; i < nbytes; i += blockDim.x) s[i] = g[i];
	}
}
__device__ __forceinline__ void rb_s2g(uint8_t* g, const uint8_t* s, int nbytes) {
	if ((reinterpret_cast<uintptr_t>(g) & 15u) == 0) {
		int nv = nbytes >> 4;
		for (int i = threadIdx.x; i < nv; i += blockDim.x)
			rb_st_stream(reinterpret_cast<uint4*>(g) + i, reinterpret_cast<const uint4*>(s)[i], RB_STORE_CS);
		for (int i = (nv << 4) + threadIdx.x; i < nbytes; i += blockDim.x) g[i] = s[i];
	} else if ((reinterpret_cast<uintptr_t>(g) & 3u) == 0) {
		int nv = nbytes >> 2;
		for (int i = threadIdx.x; i < nv; i += blockDim.x)
			reinterpret_cast<uint32_t*>(g)[i] = reinterpret_cast<const uint32_t*>(s)[i];
		for (int i = (nv << 2) + threadIdx.x; i < nbytes; i += blockDim.x) g[i] = s[i];
	} else {
		for (int i = threadIdx.x; i < nbytes; i += blockDim.x) g[i] = s[i];
	}
}
