// Shared layout constants: the "flat" fp32 parameter buffer (include/fnerf.h) and the "packed"
// device blob the kernels read.  Network shape: SURVEY.md A.4 (8x256, skip after layer 4,
// view branch 283->128->3) and A.8 (layer 5 widened by a 256-d garment code).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FN_HD __host__ __device__ __forceinline__
#else
#define FN_HD inline
#endif

namespace fnerf {

constexpr int kW = 256;          // trunk width
constexpr int kPE = 63;          // 3 + 6*10
constexpr int kPED = 27;         // 3 + 6*4
constexpr int kLX = 10, kLD = 4; // octaves
constexpr int kCond = 256;       // garment code width
constexpr int kWV = 128;         // view-branch width
constexpr int kNumLayers = 12;   // pts 0..7, alpha, feature, views, rgb

// ---- flat fp32 layout ------------------------------------------------------------------------
struct LayerDim { int out, in; };
FN_HD LayerDim layer_dim(int l, int cond) {
  if (l == 0) return {kW, kPE};
  if (l == 5) return {kW, kPE + (cond ? kCond : 0) + kW};
  if (l < 8) return {kW, kW};
  if (l == 8) return {1, kW};          // alpha_linear
  if (l == 9) return {kW, kW};         // feature_linear
  if (l == 10) return {kWV, kW + kPED}; // views_linears.0
  return {3, kWV};                     // rgb_linear
}
FN_HD int64_t flat_weight_offset(int l, int cond) {
  int64_t off = 0;
  for (int i = 0; i < l; ++i) { LayerDim d = layer_dim(i, cond); off += (int64_t)d.out * d.in + d.out; }
  return off;
}
FN_HD int64_t flat_bias_offset(int l, int cond) {
  LayerDim d = layer_dim(l, cond);
  return flat_weight_offset(l, cond) + (int64_t)d.out * d.in;
}
FN_HD int64_t flat_count(int cond) { return flat_weight_offset(kNumLayers, cond); }

// ---- packed blob -----------------------------------------------------------------------------
// Section A: bf16 UMMA chunk stream, consumption order.  A chunk is ROWS x 64 bf16, K-major,
// 128-byte swizzle (byte (r, c16) at r*128 + ((c16 ^ (r & 7)) << 4)), i.e. the exact shared-memory
// image one bulk-TMA copy drops into a pipeline stage.  Chunk kinds (chunk_desc):
//   TRUNK  W[:, base + 64*kb ..]                      A operand = activation K-block kb
//   XYZ    W[:, 0:63] (+ zero column)                 A operand = xyz-encoding tile      (layers 0, 5)
//   BIAS   one K = 16 step, 8 KB, stored MN-major (N contiguous: 4 groups of 64 outputs x 16 K rows x 128 B,
//          128-byte swizzle; byte (n, k) at (n/64)*2048 + k*128 + ((((n%64)/8) ^ (k&7)) << 4) + (n%8)*2):
//          zeros except K row 11 = bf16(b), row 12 = bf16(b - bf16(b)).
//                                                     A operand = K-step 1 (cols 16..31) of the
//          direction-encoding tile, whose cols 27, 28 hold 1.0 -> the fp32 accumulator receives the
//          bias to ~16 mantissa bits and the epilogue needs no bias add.  (A K-major BIAS chunk would be a
//          full 32 KB of which one K-step is read: 20 % of the weight stream.)
//   DIR    Wv[:, 256:283] in cols 0..26, view-layer bias hi/lo in cols 27, 28 (view layer only)
// Order: L0 {XYZ,BIAS}; L1..L4 {k0..k3,BIAS}; L5 {XYZ,k0..k3,BIAS}; L6,L7 {k0..k3,BIAS};
// feature {k0..k3,BIAS}; views {k0..k3 (128 rows), DIR (128 rows)}.
constexpr int kChunkK = 64;
constexpr int kNumChunks = 48;
constexpr int kFirstViewChunk = 43;                 // chunks 43..47: the view layer's 128-row chunks
constexpr int kNumBiasChunks = 9;
constexpr int kBigChunkBytes = 256 * kChunkK * 2;   // 32768 (also the pipeline stage size)
constexpr int kSmallChunkBytes = 128 * kChunkK * 2; // 16384
constexpr int kBiasChunkBytes = 256 * 16 * 2;       // 8192
constexpr int kBiasColHi = 27, kBiasColLo = 28;     // inside the direction tile / DIR chunk; K rows 11, 12 of a BIAS chunk
constexpr int64_t kSecABytes = (int64_t)(kFirstViewChunk - kNumBiasChunks) * kBigChunkBytes + (int64_t)kNumBiasChunks * kBiasChunkBytes +
                               (int64_t)(kNumChunks - kFirstViewChunk) * kSmallChunkBytes;

enum { CHUNK_TRUNK = 0, CHUNK_XYZ = 1, CHUNK_BIAS = 2, CHUNK_DIR = 3 };
struct ChunkDesc { int layer; int kind; int kb; };   // layer = flat layer id (0..7 trunk, 9 feature, 10 views)
FN_HD constexpr ChunkDesc chunk_desc(int c) {
  if (c == 0) return {0, CHUNK_XYZ, 0};
  if (c == 1) return {0, CHUNK_BIAS, 0};
  if (c <= 21) return {1 + (c - 2) / 5, (c - 2) % 5 == 4 ? CHUNK_BIAS : CHUNK_TRUNK, (c - 2) % 5};
  if (c == 22) return {5, CHUNK_XYZ, 0};
  if (c <= 26) return {5, CHUNK_TRUNK, c - 23};
  if (c == 27) return {5, CHUNK_BIAS, 0};
  if (c <= 37) return {6 + (c - 28) / 5, (c - 28) % 5 == 4 ? CHUNK_BIAS : CHUNK_TRUNK, (c - 28) % 5};
  if (c <= 42) return {9, c - 38 == 4 ? CHUNK_BIAS : CHUNK_TRUNK, c - 38};
  if (c <= 46) return {10, CHUNK_TRUNK, c - 43};
  return {10, CHUNK_DIR, 0};
}

FN_HD constexpr int chunk_bytes(int c) {
  return c >= kFirstViewChunk ? kSmallChunkBytes : (chunk_desc(c).kind == CHUNK_BIAS ? kBiasChunkBytes : kBigChunkBytes);
}
FN_HD constexpr int64_t chunk_offset(int c) {
  int64_t off = 0;
  for (int i = 0; i < c; ++i) off += chunk_bytes(i);
  return off;
}
// byte offset of element (output n, K row k < 16) inside a BIAS chunk
FN_HD uint32_t bias_chunk_offset(uint32_t n, uint32_t k) {
  return (n >> 6) * 2048u + k * 128u + ((((n & 63u) >> 3) ^ (k & 7u)) << 4) + (n & 7u) * 2u;
}

// Section B: fp32 "aux" (biases and the two tiny heads), offsets in floats.
constexpr int kAuxBiasPts = 0;        // 8 x 256
constexpr int kAuxBiasFeat = 2048;    // 256
constexpr int kAuxBiasViews = 2304;   // 128
constexpr int kAuxWAlpha = 2432;      // 256
constexpr int kAuxBAlpha = 2688;      // 1 (+3 pad)
constexpr int kAuxWRgb = 2692;        // 3 x 128
constexpr int kAuxBRgb = 3076;        // 3 (+pad)
constexpr int kAuxFloats = 3088;
constexpr int64_t kSecBOffset = kSecABytes;
constexpr int64_t kSecBBytes = (int64_t)kAuxFloats * 4;

// Section C: fp32 K-major (transposed) weights for the SIMT path: layer l stored as [in][out].
// Order: pts 0..7, feature, views.  (alpha / rgb live in aux.)
constexpr int64_t kSecCOffset = kSecBOffset + kSecBBytes;
FN_HD int simt_layer_id(int j) { return j < 8 ? j : (j == 8 ? 9 : 10); }  // j-th SIMT layer -> flat layer
FN_HD int64_t simt_offset_floats(int j, int cond) {
  int64_t off = 0;
  for (int i = 0; i < j; ++i) { LayerDim d = layer_dim(simt_layer_id(i), cond); off += (int64_t)d.out * d.in; }
  return off;
}
// Section E: TRANSPOSED bf16 chunk stream for the tensor-core dgrad chain (mlp_dgrad_tc.cu), consumption
// order: view layer (K-blocks 0,1 of its 128 outputs), feature layer (0..3), layers 7..1 (0..3 each).
// Chunk (layer l, kb) is 256 x 64, K-major, 128-byte swizzle: row k = INPUT feature (trunk block of
// layer 5: W5[:, hoff + k]), column j = OUTPUT feature 64*kb + j.
constexpr int kNumChunksT = 34;
constexpr int kChunkTBytes = 256 * 64 * 2;
FN_HD int64_t sec_e_offset(int cond) { return (kSecCOffset + simt_offset_floats(10, cond) * 4 + 1023) / 1024 * 1024; }
struct ChunkTDesc { int layer; int kb; };
FN_HD ChunkTDesc chunk_t_desc(int c) {
  if (c < 2) return {10, c};
  if (c < 6) return {9, c - 2};
  return {7 - (c - 6) / 4, (c - 6) % 4};
}
FN_HD int64_t packed_bytes(int cond) { return sec_e_offset(cond) + (int64_t)kNumChunksT * kChunkTBytes; }

// ---- training tape (bf16 tensor-core backward) ----------------------------------------------------
// Per 128-sample tile, a sequence of 16 KB K-block images: 128 rows (samples) x 64 bf16 columns with the
// 128-byte swizzle, i.e. the forward kernel's shared-memory bytes verbatim, so the backward kernels load
// them with one bulk copy each.  Forward slots: xyz encoding, direction encoding (+ ones in cols 27, 28),
// H0..H7 (4 K-blocks each), feature (4), view-layer output HV (2).
constexpr int kTapeSlotPe = 0, kTapeSlotPed = 1, kTapeSlotH = 2, kTapeSlotFeat = 2 + 32, kTapeSlotHv = 2 + 36;
constexpr int kTapeFwdSlots = 40;
// Backward slots (written by the dgrad chain): dZv (2), dFEAT (4), dZ7 .. dZ0 (4 each).
constexpr int kTapeBwdSlotZv = 0, kTapeBwdSlotFeat = 2, kTapeBwdSlotZ = 6;   // dZ_l at kTapeBwdSlotZ + 4*(7-l)
constexpr int kTapeBwdSlotG = 38;      // g_raw as a bf16 image: columns 0..2 = d/d rgb_raw, 3 = d/d sigma_raw (chunk 0 only is defined)
constexpr int kTapeBwdSlots = 39;
// ReLU bitmasks of the forward activations: per tile 68 units of 32 columns (H0..H7: 8 units each, HV: 4), per
// unit 128 rows x u32.  Bit i (i < 16) = column 2i of the unit, bit 16+i = column 2i+1 (set = activation > 0).
constexpr int kMaskUnits = 68, kMaskUnitHv = 64;
constexpr int kMaskTileBytes = kMaskUnits * 128 * 4;

// A.8: row of the [C,256] code table a ray reads.  Indices coming from the caller (cond_index) are clamped to the
// table, so that a bad view id can never read out of bounds (the Python layer validates and raises first).
FN_HD int64_t cond_row(const int32_t* cond_index, int64_t C, int64_t ray) {
  int64_t row = cond_index ? (int64_t)cond_index[ray] : (C == 1 ? 0 : ray);
  return row < 0 ? 0 : (row >= C ? C - 1 : row);
}

// swizzled byte offset of element (row r, column k in [0,64)) inside a chunk / activation K-block
FN_HD uint32_t sw128_offset(uint32_t r, uint32_t k) {
  return r * 128u + ((((k >> 3) ^ (r & 7u)) << 4) | ((k & 7u) << 1));
}

}  // namespace fnerf
