// Streamed BEV pool forward (bev_stream.cu): interface used by bev.cu.  Internal header.
#pragma once
#include "common.cuh"

namespace muvo {

#ifndef MUVO_STREAM_CHUNK_LOG2
#define MUVO_STREAM_CHUNK_LOG2 11
#endif
constexpr int kStreamPosBits = MUVO_STREAM_CHUNK_LOG2;   // list entry = cell << kStreamPosBits | position inside the chunk
constexpr int kStreamChunk = 1 << kStreamPosBits;        // frustum points per chunk (one bulk copy per channel)

// bytes of the two workspace regions the streamed pool needs on top of carve_bev's
size_t stream_lists_bytes(int B, int64_t n_pts);
size_t stream_steps_bytes(int B, int64_t n_pts);

// true when the streamed kernel can serve this call (layout, alignment, shared memory, enough work items)
bool pool_stream_eligible(int elem_bytes, const void* x, int64_t sb, int64_t sp, int64_t sc, int B, int64_t n_pts, int C, int n_cells);

// the mask-independent part of the per-chunk lists (sorted keys + counts): plan [stream_lists_bytes], plan_n [stream_steps_bytes]
int pool_stream_plan(const int32_t* cell0, int B, int64_t n_pts, int n_cells, uint32_t* plan, uint32_t* plan_n, cudaStream_t st);

// cell0 [B, n_pts] int32 (-1 = dropped), mask [B, n_pts] uint8 or nullptr, cell_out [B, n_pts] (cell0 with the mask folded in)
// or nullptr; plan / plan_n from pool_stream_plan(cell0) or nullptr (then the lists are sorted per call); out [B, C, n_cells]
// float32 fully written
int pool_stream_fwd(const void* x, int32_t x_dtype, int64_t sb, int64_t sc, const int32_t* cell0, const uint8_t* mask,
                    int32_t* cell_out, const uint32_t* plan, const uint32_t* plan_n, int B, int64_t n_pts, int C, int n_cells,
                    float* out, uint32_t* lists, uint32_t* steps, cudaStream_t st);

// Streamed backward (same layout conditions, applied to the gradient tensor): `lists` / `steps` are the ones the streamed
// forward left in its workspace for the same (cell0, mask)
bool pool_bwd_stream_eligible(int elem_bytes, const void* gx, int64_t sb, int64_t sp, int64_t sc, int B, int64_t n_pts, int C, int n_cells);
int pool_stream_bwd(const float* gout, const uint32_t* lists, const uint32_t* steps, int B, int64_t n_pts, int C, int n_cells, void* gx,
                    int32_t gx_dtype, int64_t sb, int64_t sc, cudaStream_t st);

}  // namespace muvo
