// Host-side generation of the move tables and their upload to the current device.
// Row a1 of SURVEY.md 8(a): the tables are derived here from the move definitions, not retyped.
#pragma once
#include "rb_common.cuh"
#include <stdlib.h>

namespace rbt {

// Positive-direction 4-cycles of corner / edge positions for F, B, T, D, L, R and the orientation rule
// (reference: librubiks/cube/maps.py:74-98 `Actions`).
static const int kCornerCycle[6][4] = {{0, 1, 2, 3}, {4, 7, 6, 5}, {0, 3, 7, 4}, {1, 5, 6, 2}, {0, 4, 5, 1}, {7, 3, 2, 6}};
static const int kEdgeCycle[6][4] = {{0, 1, 2, 3}, {8, 11, 10, 9}, {0, 7, 8, 4}, {2, 5, 10, 6}, {1, 4, 9, 5}, {3, 6, 11, 7}};
static const int kCornerStatic[6] = {0, 0, 1, 1, 2, 2};   // this orientation is kept, the other two swap
static const bool kEdgeFlips[6] = {false, false, true, true, false, false};

// 6x8x6 constants (reference: librubiks/cube/maps.py:149-156, cube.py:311-326).
static const int kNeighbors686[6][4] = {{4, 3, 5, 2}, {3, 4, 2, 5}, {0, 5, 1, 4}, {5, 0, 4, 1}, {2, 1, 3, 0}, {1, 2, 0, 3}};
static const int kAdjacent[12] = {6, 7, 0, 2, 3, 4, 4, 5, 6, 0, 1, 2};

struct Tables {
	int8_t delta[2][6][2][24];    // get_tensor_map layout: [dir 0=neg,1=pos][face][kind 0=corner,1=edge][value]
	uint8_t lut[12][2][24];       // direct form per action index
	uint8_t lut_padded[12 * 2 * 32];
	uint8_t perm686[12][48];      // gather table over the 48 sticker slots
	uint8_t solved2024[32];
	uint8_t solved686[288];
};

static void build(Tables& t) {
	memset(&t, 0, sizeof(t));
	// maps.py:107-145
	for (int f = 0; f < 6; ++f)
		for (int j = 0; j < 4; ++j) {
			for (int k = 0; k < 3; ++k) {
				int nk = (k == kCornerStatic[f]) ? k : 3 - kCornerStatic[f] - k;
				int src = 3 * kCornerCycle[f][j] + k, dst = 3 * kCornerCycle[f][(j + 1) & 3] + nk;
				t.delta[1][f][0][src] = (int8_t)(dst - src);
				t.delta[0][f][0][dst] = (int8_t)(src - dst);
			}
			for (int k = 0; k < 2; ++k) {
				int nk = kEdgeFlips[f] ? 1 - k : k;
				int src = 2 * kEdgeCycle[f][j] + k, dst = 2 * kEdgeCycle[f][(j + 1) & 3] + nk;
				t.delta[1][f][1][src] = (int8_t)(dst - src);
				t.delta[0][f][1][dst] = (int8_t)(src - dst);
			}
		}
	for (int a = 0; a < 12; ++a) {
		int f = a / 2, d = 1 - a % 2;                 // cube.py:33-35
		for (int kind = 0; kind < 2; ++kind)
			for (int s = 0; s < 32; ++s) {
				uint8_t v = (uint8_t)(s < 24 ? s + t.delta[d][f][kind][s] : s);
				if (s < 24) t.lut[a][kind][s] = v;
				t.lut_padded[(a * 2 + kind) * 32 + s] = v;
			}
	}
	// cube.py:330-347 as a gather over slot labels: perm[a][dst] = src.
	int rolled[12];
	for (int i = 0; i < 12; ++i) rolled[(i + 3) % 12] = kAdjacent[i];
	for (int a = 0; a < 12; ++a) {
		int f = a / 2, d = 1 - a % 2;
		uint8_t* p = t.perm686[a];
		for (int s = 0; s < 48; ++s) p[s] = (uint8_t)s;
		for (int i = 0; i < 8; ++i)                   // the turned face's ring shifts by two
			p[f * 8 + i] = (uint8_t)(f * 8 + (d ? (i + 6) % 8 : (i + 2) % 8));
		for (int i = 0; i < 12; ++i) {                // 12 adjacent stickers move one neighbour face along
			int blk = i / 3, prev = (blk + 3) % 4;
			if (d) p[kNeighbors686[f][blk] * 8 + kAdjacent[i]] = (uint8_t)(kNeighbors686[f][prev] * 8 + rolled[i]);
			else   p[kNeighbors686[f][prev] * 8 + rolled[i]] = (uint8_t)(kNeighbors686[f][blk] * 8 + kAdjacent[i]);
		}
	}
	// cube.py:58-71
	for (int i = 0; i < 8; ++i) t.solved2024[i] = (uint8_t)(3 * i);
	for (int i = 0; i < 12; ++i) t.solved2024[8 + i] = (uint8_t)(2 * i);
	for (int f = 0; f < 6; ++f)
		for (int p = 0; p < 8; ++p) t.solved686[(f * 8 + p) * 6 + f] = 1;
}

static const Tables& host() {
	static Tables t;
	static std::once_flag once;
	std::call_once(once, [] { build(t); });
	return t;
}

// Upload to the current device once.
static int ensure_device() {
	static std::mutex mu;
	static bool done[64] = {};
	int dev = 0;
	RB_CUDA(cudaGetDevice(&dev));
	if (dev < 0 || dev >= 64) return rb_fail(RB_ERR_BAD_ARG, "device ordinal out of range%s%s");
	std::lock_guard<std::mutex> lock(mu);
	if (done[dev]) return RB_OK;
	const Tables& t = host();
	RB_CUDA(cudaMemcpyToSymbol(g_lut2024, t.lut_padded, sizeof(t.lut_padded)));
	RB_CUDA(cudaMemcpyToSymbol(c_lut2024, t.lut_padded, sizeof(t.lut_padded)));
	RB_CUDA(cudaMemcpyToSymbol(g_perm686, t.perm686, sizeof(t.perm686)));
	RB_CUDA(cudaMemcpyToSymbol(g_solved2024, t.solved2024, sizeof(t.solved2024)));
	RB_CUDA(cudaMemcpyToSymbol(g_solved686, t.solved686, sizeof(t.solved686)));
	done[dev] = true;
	return RB_OK;
}

}  // namespace rbt
