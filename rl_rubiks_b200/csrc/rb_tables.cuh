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
	// Sticker view of the cubies (derived below by running both representations side by side): the 6x8x6 slot of sticker k
	// of corner cubie c / edge cubie e at home, and its slot when the cubie's 20x24 value is v.  Used to render a 6x8x6
	// state from the 20x24 state of the same move sequence (rb686 fast scramble).
	uint8_t corner_home[8][3], corner_dst[8][24][3];
	uint8_t edge_home[12][2], edge_dst[12][24][2];
	bool stickers_ok;
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

// Follows one cubie through all 24 (position, orientation) values with the 20x24 LUT while moving its stickers with the
// 6x8x6 slot permutations; `dst[v][k]` = slot of sticker k when the cubie's value is v.  Returns false on any inconsistency
// between the two representations.
template <int K>
static bool follow_cubie(const Tables& t, int kind, int home_value, const uint8_t (&home)[K], uint8_t (*dst)[K]) {
	bool seen[24] = {};
	int queue[24], head = 0, tail = 0;
	for (int k = 0; k < K; ++k) dst[home_value][k] = home[k];
	seen[home_value] = true;
	queue[tail++] = home_value;
	while (head < tail) {
		const int v = queue[head++];
		for (int a = 0; a < 12; ++a) {
			const int v2 = t.lut[a][kind][v];
			uint8_t moved[K];
			for (int k = 0; k < K; ++k) {
				int to = -1;
				for (int x = 0; x < 48; ++x) if (t.perm686[a][x] == dst[v][k]) to = x;     // new[x] = old[perm[x]]
				if (to < 0) return false;
				moved[k] = (uint8_t)to;
			}
			if (!seen[v2]) {
				seen[v2] = true;
				for (int k = 0; k < K; ++k) dst[v2][k] = moved[k];
				queue[tail++] = v2;
			} else {
				for (int k = 0; k < K; ++k) if (dst[v2][k] != moved[k]) return false;
			}
		}
	}
	return tail == 24;
}

static void build_stickers(Tables& t) {
	t.stickers_ok = false;
	int moved_by[48] = {};                                   // bitmask of faces whose turn moves the slot
	for (int f = 0; f < 6; ++f)
		for (int x = 0; x < 48; ++x) if (t.perm686[2 * f][x] != x) moved_by[x] |= 1 << f;
	// corner position p: its three facelets are the slots moved by exactly the three faces that move the position; the
	// 20x24 orientation is the axis (F/B, T/D, L/R) the tracked sticker faces (maps.py:128), so sticker k sits on axis k at home
	uint8_t corner_slot[8][3];
	for (int p = 0; p < 8; ++p) {
		int faces = 0, found = 0;
		for (int f = 0; f < 6; ++f) if (t.lut[2 * f][0][3 * p] != 3 * p) faces |= 1 << f;
		for (int x = 0; x < 48; ++x) if (moved_by[x] == faces) { corner_slot[p][(x / 8) / 2] = (uint8_t)x; ++found; }
		if (found != 3) return;
	}
	for (int c = 0; c < 8; ++c) {
		for (int k = 0; k < 3; ++k) t.corner_home[c][k] = corner_slot[c][k];
		if (!follow_cubie<3>(t, 0, 3 * c, t.corner_home[c], t.corner_dst[c])) return;
		for (int v = 0; v < 24; ++v) if (t.corner_dst[c][v][0] != corner_slot[v / 3][v % 3]) return;   // tracked sticker on axis o
	}
	// edges: which of the two facelets carries the tracked sticker is fixed up to one global swap that the gather does not see;
	// cubie 0 defines it for every (position, orientation), the other cubies must agree
	uint8_t edge_slot[12][2];
	for (int p = 0; p < 12; ++p) {
		int faces = 0, found = 0;
		for (int f = 0; f < 6; ++f) if (t.lut[2 * f][1][2 * p] != 2 * p) faces |= 1 << f;
		for (int x = 0; x < 48; ++x) if (moved_by[x] == faces) { if (found < 2) edge_slot[p][found] = (uint8_t)x; ++found; }
		if (found != 2) return;
	}
	t.edge_home[0][0] = edge_slot[0][0];
	t.edge_home[0][1] = edge_slot[0][1];
	if (!follow_cubie<2>(t, 1, 0, t.edge_home[0], t.edge_dst[0])) return;
	for (int e = 1; e < 12; ++e) {
		t.edge_home[e][0] = t.edge_dst[0][2 * e][0];
		t.edge_home[e][1] = t.edge_dst[0][2 * e][1];
		if (!follow_cubie<2>(t, 1, 2 * e, t.edge_home[e], t.edge_dst[e])) return;
		for (int v = 0; v < 24; ++v) if (t.edge_dst[e][v][0] != t.edge_dst[0][v][0]) return;
	}
	t.stickers_ok = true;
}

static const Tables& host() {
	static Tables t;
	static std::once_flag once;
	std::call_once(once, [] { build(t); build_stickers(t); });
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
	{
		// sticker tables, flat: [cubie 0..19][value 0..23][k 0..2] destination slots, [cubie][k] home slots (k = 2 unused for edges)
		static uint8_t dst[20 * 24 * 3 + 20 * 3 + 4];
		for (int c = 0; c < 20; ++c)
			for (int k = 0; k < 3; ++k) {
				const bool corner = c < 8, used = corner || k < 2;
				dst[20 * 24 * 3 + c * 3 + k] = used ? (corner ? t.corner_home[c][k] : t.edge_home[c - 8][k]) : 0;
				for (int v = 0; v < 24; ++v) dst[(c * 24 + v) * 3 + k] = used ? (corner ? t.corner_dst[c][v][k] : t.edge_dst[c - 8][v][k]) : 0;
			}
		RB_CUDA(cudaMemcpyToSymbol(g_stickers686, dst, sizeof(dst)));
	}
	done[dev] = true;
	return RB_OK;
}

}  // namespace rbt
