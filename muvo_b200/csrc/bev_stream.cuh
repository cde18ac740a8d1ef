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

// cell0 [B, n_pts] int32 (-1 = dropped), mask [B, n_pts] uint8 or nullptr, cell_out [B, n_pts] (cell0 with the mask folded in)
// or nullptr; out [B, C, n_cells] float32 fully written
int pool_stream_fwd(const void* x, int32_t x_dtype, int64_t sb, int64_t sc, const int32_t* cell0, const uint8_t* mask,
                    int32_t* cell_out, int B, int64_t n_pts, int C, int n_cells, float* out, uint32_t* lists, uint32_t* steps,
                    cudaStream_t st);

}  // namespace muvo
