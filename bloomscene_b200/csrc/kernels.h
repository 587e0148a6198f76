// Internal launcher interface between api.cu and the kernel translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace brs {

void count_launch();

// ---- preprocess.cu ------------------------------------------------------------------------------
struct PreprocessArgs {
	int P, D, M;
	const float* means3D;
	const float* scales;
	float scale_modifier;
	const float* rotations;
	const float* opacities;
	const float* shs;
	const float* cov3D_precomp;
	const float* colors_precomp;
	const float* viewmatrix;
	const float* projmatrix;
	const float* campos;
	int W, H;
	float tan_fovx, tan_fovy, focal_x, focal_y;
	uint32_t grid_x, grid_y;
	int prefiltered;
	// outputs
	int* radii;
	float4* records;      // [3P]
	uint32_t* depth_key;  // [P]
	uint2* rect;          // [P]
	uint32_t* total_tiles; // [5], pre-zeroed: R = tile instances, R1 = supertile instances, max(~depth bits),
	                       // max(depth bits) over visible, V = number of visible Gaussians
};
cudaError_t launch_preprocess(const PreprocessArgs& a, cudaStream_t stream);

struct FilterArgs {
	int P;
	const float* means3D;
	const float* scales;
	int scales_stride;
	float scale_modifier;
	const float* rotations;
	const float* cov3D_precomp;
	const float* viewmatrix;
	const float* projmatrix;
	int W, H;
	float tan_fovx, tan_fovy, focal_x, focal_y;
	uint32_t grid_x, grid_y;
	int prefiltered;
	int* radii;
};
cudaError_t launch_filter(const FilterArgs& a, cudaStream_t stream);
cudaError_t launch_check_frustum(int P, const float* means3D, const float* viewmatrix, uint8_t* present,
                                 cudaStream_t stream);

// ---- binning.cu ---------------------------------------------------------------------------------
// Stable LSD radix sort of (u32,u32) pairs, onesweep-style (single pass per digit with decoupled
// look-back).  Scratch layout is private; size from sort_scratch_bytes(n).
size_t sort_scratch_bytes(size_t n);
// Sorts on key bits [begin_bit, end_bit).  vals_in == nullptr -> iota.  The result is written to
// keys_out/vals_out; keys_in/vals_in are not modified.  tmp buffers for ping-pong live in scratch.
cudaError_t sort_pairs(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out,
                       size_t n, int begin_bit, int end_bit, void* scratch, cudaStream_t stream);
// One stable pass on the digit ((key - bias) >> shift) & ((1 << bits) - 1), bits <= 8; in != out.
// With `drop`, keys equal to `drop_key` are left out: the output then holds only the other keys
// (compacted, still stable) and the caller continues with the smaller count.
cudaError_t sort_pass(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, size_t n,
                      uint32_t bias, int shift, int bits, void* scratch, cudaStream_t stream, bool drop = false,
                      uint32_t drop_key = 0);
// The ping-pong buffers inside a sort scratch area (n keys + n values).
void sort_tmp_buffers(void* scratch, size_t n, uint32_t** tmp_keys, uint32_t** tmp_vals);

// Fused exclusive scan of per-Gaussian cell counts (in depth order) + emission of (cell, id)
// instances, a cell being (1 << shift)^2 tiles and grid_x the number of cells per row.
// `order` = Gaussian ids sorted by depth; `rect` packed tile rectangles.
size_t emit_scratch_bytes(size_t P);
cudaError_t launch_emit(const uint32_t* order, const uint2* rect, size_t P, uint32_t shift, uint32_t grid_x,
                        uint32_t* cell_keys, uint32_t* inst_ids, size_t n_instances, void* scratch, cudaStream_t stream);

// ranges[key] = [first, last+1) in a sorted key list; `ranges` must be pre-zeroed.
cudaError_t launch_tile_ranges(const uint32_t* sorted_keys, size_t n, uint2* ranges, cudaStream_t stream);

// Fine level of the binning: per-supertile lists (sorted_coarse_keys / coarse_list, R1 entries in
// (supertile, depth, id) order) -> point_list (R ids in (tile, depth, id) order) and tile ranges.
size_t fine_scratch_bytes(size_t R1, uint32_t grid_x, uint32_t grid_y);
cudaError_t launch_fine_binning(const uint32_t* sorted_coarse_keys, const uint32_t* coarse_list, size_t R1,
                                const uint2* rect, uint32_t grid_x, uint32_t grid_y, uint32_t* point_list,
                                uint2* ranges, void* scratch, cudaStream_t stream);

// ---- blend_fwd.cu -------------------------------------------------------------------------------
struct BlendFwdArgs {
	float exp_c_scale, exp_c_252; // constants of expf_exact (common.cuh), filled by the launcher
	const uint2* ranges;
	const uint32_t* point_list;
	const float4* records;
	const float* bg;
	int W, H;
	uint32_t grid_x, grid_y;
	float* final_T;
	uint32_t* n_contrib;
	float* out_color;
	float* out_depth;
};
cudaError_t launch_blend_forward(const BlendFwdArgs& a, cudaStream_t stream);

// ---- blend_bwd.cu -------------------------------------------------------------------------------
struct BlendBwdArgs {
	float exp_c_scale, exp_c_252; // constants of expf_exact (common.cuh), filled by the launcher
	const uint2* ranges;
	const uint32_t* point_list;
	const float4* records;
	const float* bg;
	int W, H;
	uint32_t grid_x, grid_y;
	const float* final_T;
	const uint32_t* n_contrib;
	const float* dL_dpixels; // [3,H,W]
	const float* dL_ddepth;  // [H,W]; with out_depth: opt-in depth gradient (brs_grads.depth_gradient), else nullptr
	const float* out_depth;  // [H,W] the forward's depth output
	float* accum;            // [P][12] fp32, pre-zeroed: {dmean2D.x, .y, dconic.x, .y, .w, dopacity, dcol r,g,b, dz, pad x2}
};
cudaError_t launch_blend_backward(const BlendBwdArgs& a, cudaStream_t stream);

// ---- preprocess_bwd.cu --------------------------------------------------------------------------
struct PreprocessBwdArgs {
	int P, D, M;
	const float* means3D;
	const int* radii;
	const float* shs;
	const float* scales;
	const float* rotations;
	float scale_modifier;
	const float* cov3D_precomp;
	const float* viewmatrix;
	const float* projmatrix;
	const float* campos;
	int W, H;
	float tan_fovx, tan_fovy, focal_x, focal_y;
	const float* accum; // [P][12] from the blend backward
	int accumulate;     // 0: every output element is written; 1: see brs_grads.accumulate
	int depth_gradient; // 1: accum slot 9 holds dL_dz (view-space depth) and feeds dL_dmeans3D
	int fact_offset;    // (set by the launcher) float offset of the basis-factor rows in dynamic shared memory
	// outputs
	float* dL_dmeans2D;   // [P,3]
	float* dL_dcolors;    // [P,3]
	float* dL_dopacity;   // [P]
	float* dL_dmeans3D;   // [P,3]
	float* dL_dcov3D;     // [P,6]
	float* dL_dsh;        // [P,M,3] or nullptr
	float* dL_dscales;    // [P,3]
	float* dL_drotations; // [P,4]
};
cudaError_t launch_preprocess_backward(const PreprocessBwdArgs& a, cudaStream_t stream);

// ---- measure.cu (bench / profiling only) ----------------------------------------------------------
cudaError_t launch_count_pairs(const BlendFwdArgs& a, unsigned long long* out, cudaStream_t stream);
double probe_fp32_tflops(cudaStream_t stream);

constexpr int ACCUM_STRIDE = 12; // floats per Gaussian in the blend-backward accumulator

} // namespace brs
