// C-ABI of libbloomrast (see include/bloomrast.h): stage orchestration, private state layout,
// error handling.  Replaces the reference driver CudaRasterizer::Rasterizer::{forward, backward,
// visible_filter, markVisible} (cuda_rasterizer/rasterizer_impl.cu:141-504) and its chunk
// sub-allocator (rasterizer_impl.h:21-73, rasterizer_impl.cu:155-194).
//
// Stage order of brs_forward (all on the caller's stream):
//   memset(header) -> preprocess -> [async copy of R, R1 and the depth-key range to pinned host + event]
//   -> depth sort pass 1 (low 8 key bits of P keys; needs none of them, so the GPU stays busy while the
//      host waits for the event) -> host: wait, allocate binning/scratch, choose the remaining digit
//      widths from the key range -> depth sort passes 2..k -> coarse binning (scan + emit of supertile
//      instances, one radix pass on the supertile id) -> fine binning (count, scan, scatter:
//      point_list and tile ranges) -> blend.
// brs_backward: memset(accumulator) -> blend backward -> fused preprocess backward (plain or
// accumulate mode, see brs_grads).  No host sync.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "kernels.h"

namespace brs {

namespace {

thread_local long long t_launches = 0;
thread_local int t_last_cuda_error = 0;

struct HostSlot { // per thread: pinned word for the R readback + its event
	int device = -1;
	uint32_t* pinned = nullptr;
	cudaEvent_t event = nullptr;
	~HostSlot() // host threads come and go (one per lane of a view batch): give the pinned word and the event back
	{
		if (pinned != nullptr)
			cudaFreeHost(pinned);
		if (event != nullptr)
			cudaEventDestroy(event);
	}
};
thread_local HostSlot t_slot;

cudaError_t ensure_slot()
{
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess)
		return e;
	if (t_slot.pinned != nullptr && t_slot.device == dev)
		return cudaSuccess;
	if (t_slot.pinned != nullptr) {
		cudaFreeHost(t_slot.pinned);
		cudaEventDestroy(t_slot.event);
		t_slot.pinned = nullptr;
		t_slot.event = nullptr;
		t_slot.device = -1;
	}
	e = cudaHostAlloc(reinterpret_cast<void**>(&t_slot.pinned), 64, cudaHostAllocDefault);
	if (e != cudaSuccess)
		return e;
	e = cudaEventCreateWithFlags(&t_slot.event, cudaEventDisableTiming);
	if (e != cudaSuccess)
		return e;
	t_slot.device = dev;
	return cudaSuccess;
}

// Optional per-stage CUDA-event timing (brs_stage_timing / brs_stage_times).
struct StageTimer {
	bool enabled = false;
	struct Pending { int stage; cudaEvent_t a, b; };
	std::vector<Pending> pending;
	std::vector<cudaEvent_t> pool;
	~StageTimer()
	{
		for (auto& p : pending) {
			cudaEventDestroy(p.a);
			cudaEventDestroy(p.b);
		}
		for (cudaEvent_t e : pool)
			cudaEventDestroy(e);
	}
	cudaEvent_t get()
	{
		if (!pool.empty()) {
			cudaEvent_t e = pool.back();
			pool.pop_back();
			return e;
		}
		cudaEvent_t e = nullptr;
		cudaEventCreate(&e);
		return e;
	}
};
thread_local StageTimer t_timer;

struct StageScope {
	int stage;
	cudaStream_t stream;
	cudaEvent_t a = nullptr;
	StageScope(int stage_, cudaStream_t s) : stage(stage_), stream(s)
	{
		if (t_timer.enabled) {
			a = t_timer.get();
			cudaEventRecord(a, stream);
		}
	}
	~StageScope()
	{
		if (a != nullptr) {
			cudaEvent_t b = t_timer.get();
			cudaEventRecord(b, stream);
			t_timer.pending.push_back({stage, a, b});
		}
	}
};

// Companion stream for the blend kernels.  The blend kernels are large, issue-bound grids; everything
// else on the path is small or bandwidth-bound.  When the caller runs several views on several streams
// (view-sharded step), a blend grid monopolises the block scheduler and the other views' small kernels
// wait for its tail.  If the caller's stream has a HIGHER priority than the device's lowest, the blend
// kernels are therefore launched on a lowest-priority companion stream, fenced by events on both
// sides: stream order as the caller sees it is unchanged, but pending blocks of the (higher-priority)
// preprocess / sort / binning kernels of other views are dispatched ahead of the blend's remaining
// blocks and run underneath it.  Callers on default-priority streams get no companion.
struct Companion {
	cudaStream_t stream = nullptr;
	cudaEvent_t before = nullptr, after = nullptr;
};
std::mutex g_companion_mutex;
std::map<std::pair<int, cudaStream_t>, Companion> g_companions;
int g_companion_enabled = 1;

// Returns false when the blend should simply run on `stream`; otherwise `out` is a copy of the entry
// (handles only: the table may be rebuilt by another thread at any time).
bool companion_for(cudaStream_t stream, Companion& out)
{
	if (!g_companion_enabled || t_timer.enabled || stream == nullptr)
		return false;
	int least = 0, greatest = 0, prio = 0, dev = 0;
	if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess || least == greatest)
		return false;
	if (cudaStreamGetPriority(stream, &prio) != cudaSuccess || prio >= least) // numerically lower = higher priority
		return false;
	if (cudaGetDevice(&dev) != cudaSuccess)
		return false;
	std::lock_guard<std::mutex> lock(g_companion_mutex);
	if (g_companions.size() >= 64 && g_companions.find(std::make_pair(dev, stream)) == g_companions.end()) {
		// callers that keep creating streams would grow the table without bound: drop it and start over
		// (destroying a stream / event with work in flight is legal, the driver releases it afterwards)
		for (auto& kv : g_companions) {
			cudaStreamDestroy(kv.second.stream);
			cudaEventDestroy(kv.second.before);
			cudaEventDestroy(kv.second.after);
		}
		g_companions.clear();
	}
	Companion& c = g_companions[std::make_pair(dev, stream)];
	if (c.stream == nullptr) {
		if (cudaStreamCreateWithPriority(&c.stream, cudaStreamNonBlocking, least) != cudaSuccess ||
		    cudaEventCreateWithFlags(&c.before, cudaEventDisableTiming) != cudaSuccess ||
		    cudaEventCreateWithFlags(&c.after, cudaEventDisableTiming) != cudaSuccess) {
			c = Companion{};
			return false;
		}
	}
	out = c;
	return true;
}

inline int fail_cuda(cudaError_t e)
{
	t_last_cuda_error = (int)e;
	return BRS_ERR_CUDA;
}

#define BRS_CUDA(expr)                                                                                                 \
	do {                                                                                                               \
		cudaError_t _e = (expr);                                                                                       \
		if (_e != cudaSuccess)                                                                                         \
			return fail_cuda(_e);                                                                                      \
	} while (0)

// reference CHECK_CUDA (auxiliary.h:166-173): only when debug, synchronise and surface errors
#define BRS_STAGE(stage_id, expr, debug, stream)                                                                       \
	do {                                                                                                               \
		StageScope _scope(stage_id, stream);                                                                           \
		BRS_CUDA(expr);                                                                                                \
		if (debug)                                                                                                     \
			BRS_CUDA(cudaStreamSynchronize(stream));                                                                   \
	} while (0)

constexpr size_t HEADER_BYTES = 256;

struct GeomLayout {
	size_t header, records, depth_key, rect, order, total;
};
GeomLayout geom_layout(size_t P)
{
	GeomLayout l{};
	size_t off = 0;
	l.header = off;
	off += HEADER_BYTES;
	l.records = off;
	off += align_up(sizeof(float4) * 3 * P, 256);
	l.depth_key = off;
	off += align_up(sizeof(uint32_t) * P, 256);
	l.rect = off;
	off += align_up(sizeof(uint2) * P, 256);
	l.order = off;
	off += align_up(sizeof(uint32_t) * P, 256);
	l.total = off;
	return l;
}

struct ImageLayout {
	size_t ranges, final_T, n_contrib, total;
};
ImageLayout image_layout(int W, int H)
{
	ImageLayout l{};
	const size_t gx = (W + TILE_X - 1) / TILE_X, gy = (H + TILE_Y - 1) / TILE_Y;
	const size_t npix = (size_t)W * H;
	size_t off = 0;
	l.ranges = off;
	off += align_up(sizeof(uint2) * gx * gy, 256);
	l.final_T = off;
	off += align_up(sizeof(float) * npix, 256);
	l.n_contrib = off;
	off += align_up(sizeof(uint32_t) * npix, 256);
	l.total = off < 256 ? 256 : off;
	return l;
}

size_t binning_bytes(size_t R) { return align_up(sizeof(uint32_t) * (R ? R : 1), 256); }

// reference getHigherMsb (rasterizer_impl.cu:35-50) yields the bit count the tile id is sorted on;
// any bit count >= ceil(log2(#tiles)) gives the same order, so use the exact one.
int tile_bits(uint32_t num_tiles)
{
	int b = 0;
	while ((1ull << b) < (unsigned long long)num_tiles)
		b++;
	return b < 1 ? 1 : b;
}

int validate_view(const brs_view* v, bool need_campos_bg)
{
	if (v == nullptr || v->viewmatrix == nullptr || v->projmatrix == nullptr)
		return BRS_ERR_INVALID_ARG;
	if (v->image_width < 0 || v->image_height < 0)
		return BRS_ERR_INVALID_ARG;
	if (v->image_width >= 65536 * TILE_X || v->image_height >= 65536 * TILE_Y)
		return BRS_ERR_UNSUPPORTED; // tile coordinates are packed into 16 bits
	if (need_campos_bg && v->bg == nullptr)
		return BRS_ERR_INVALID_ARG;
	return BRS_OK;
}

int validate_gaussians(const brs_view* v, const brs_gaussians* g)
{
	if (g == nullptr || g->P < 0)
		return BRS_ERR_INVALID_ARG;
	if (g->P == 0)
		return BRS_OK;
	if (g->means3D == nullptr || g->opacities == nullptr)
		return BRS_ERR_INVALID_ARG;
	// reference Python wrapper raises for both / neither (depth_diff_gaussian_rasterization/__init__.py:192-196)
	if ((g->shs == nullptr) == (g->colors_precomp == nullptr))
		return BRS_ERR_INVALID_ARG;
	const bool has_sr = g->scales != nullptr && g->rotations != nullptr;
	if (has_sr == (g->cov3D_precomp != nullptr))
		return BRS_ERR_INVALID_ARG;
	if (!has_sr && (g->scales != nullptr || g->rotations != nullptr))
		return BRS_ERR_INVALID_ARG;
	if (g->shs != nullptr) {
		if (v->sh_degree < 0 || v->sh_degree > 3)
			return BRS_ERR_UNSUPPORTED;
		if (v->sh_coeffs < (v->sh_degree + 1) * (v->sh_degree + 1) || v->sh_coeffs > 16)
			return BRS_ERR_INVALID_ARG;
		if (v->campos == nullptr)
			return BRS_ERR_INVALID_ARG;
	}
	return BRS_OK;
}

} // namespace

void count_launch() { t_launches++; }

} // namespace brs

using namespace brs;

extern "C" {

int brs_version(void) { return BRS_VERSION; }

long long brs_launch_count(int reset)
{
	long long v = t_launches;
	if (reset)
		t_launches = 0;
	return v;
}

int brs_last_cuda_error(void) { return t_last_cuda_error; }
const char* brs_last_cuda_error_string(void) { return cudaGetErrorString((cudaError_t)t_last_cuda_error); }

const char* brs_error_string(int status)
{
	switch (status) {
	case BRS_OK: return "ok";
	case BRS_ERR_INVALID_ARG: return "invalid argument";
	case BRS_ERR_ALLOC: return "allocator callback returned NULL";
	case BRS_ERR_CUDA: return "CUDA error (see brs_last_cuda_error_string)";
	case BRS_ERR_UNSUPPORTED: return "unsupported configuration";
	case BRS_ERR_STATE: return "forward state does not match the given sizes";
	default: return "unknown status";
	}
}

size_t brs_geom_bytes(int P) { return geom_layout(P < 0 ? 0 : (size_t)P).total; }
size_t brs_binning_bytes(int R) { return binning_bytes(R < 0 ? 0 : (size_t)R); }
size_t brs_image_bytes(int W, int H) { return image_layout(W, H).total; }
size_t brs_sort_scratch_bytes(int n) { return sort_scratch_bytes(n < 0 ? 0 : (size_t)n); }

// sorted keys (final) + first-pass keys/values + the sort's own scratch (tables + ping-pong pair)
static size_t depth_scratch_bytes(size_t P) { return 3 * align_up(sizeof(uint32_t) * P, 256) + sort_scratch_bytes(P); }
// R1 = supertile instances (what the coarse level sorts); never more than the tile instances R.
static size_t instance_scratch_bytes(size_t P, size_t R1, uint32_t grid_x, uint32_t grid_y)
{
	return 4 * align_up(sizeof(uint32_t) * R1, 256) + sort_scratch_bytes(R1) + emit_scratch_bytes(P) +
	       fine_scratch_bytes(R1, grid_x, grid_y);
}
size_t brs_forward_scratch_bytes(int P, int R, int W, int H)
{
	const uint32_t gx = W > 0 ? (W + TILE_X - 1) / TILE_X : 0, gy = H > 0 ? (H + TILE_Y - 1) / TILE_Y : 0;
	return depth_scratch_bytes(P < 0 ? 0 : P) + instance_scratch_bytes(P < 0 ? 0 : P, R < 0 ? 0 : R, gx, gy);
}
size_t brs_backward_scratch_bytes(int P) { return align_up(sizeof(float) * ACCUM_STRIDE * (P < 0 ? 0 : (size_t)P), 256); }

int brs_state_layout(int P, int R, int W, int H, brs_layout* out)
{
	if (out == nullptr || P < 0 || R < 0 || W < 0 || H < 0)
		return BRS_ERR_INVALID_ARG;
	const GeomLayout g = geom_layout(P);
	const ImageLayout im = image_layout(W, H);
	out->geom_records = g.records;
	out->geom_depth_key = g.depth_key;
	out->geom_rect = g.rect;
	out->geom_order = g.order;
	out->binning_point_list = 0;
	out->image_ranges = im.ranges;
	out->image_final_T = im.final_T;
	out->image_n_contrib = im.n_contrib;
	return BRS_OK;
}

void brs_stage_timing(int enable) { t_timer.enabled = enable != 0; }

int brs_blend_companion_stream(int enable)
{
	const int old = g_companion_enabled;
	if (enable >= 0)
		g_companion_enabled = enable != 0;
	return old;
}

int brs_stage_times(float* ms, int* calls)
{
	for (auto& p : t_timer.pending) {
		BRS_CUDA(cudaEventSynchronize(p.b));
		float t = 0.f;
		BRS_CUDA(cudaEventElapsedTime(&t, p.a, p.b));
		if (p.stage >= 0 && p.stage < BRS_NUM_STAGES) {
			if (ms)
				ms[p.stage] += t;
			if (calls)
				calls[p.stage] += 1;
		}
		t_timer.pool.push_back(p.a);
		t_timer.pool.push_back(p.b);
	}
	t_timer.pending.clear();
	return BRS_OK;
}

double brs_probe_fp32_tflops(brs_stream stream) { return probe_fp32_tflops(stream); }

int brs_count_pairs(const brs_view* view, const brs_fwd_state* state, int P, unsigned long long* out, brs_stream stream)
{
	if (view == nullptr || state == nullptr || out == nullptr || P < 0)
		return BRS_ERR_INVALID_ARG;
	const int W = view->image_width, H = view->image_height;
	const GeomLayout gl = geom_layout(P);
	const ImageLayout il = image_layout(W, H);
	if (P == 0 || state->geom == nullptr || state->image == nullptr) {
		BRS_CUDA(cudaMemsetAsync(out, 0, 3 * sizeof(unsigned long long), stream));
		return BRS_OK;
	}
	const char* geom = static_cast<const char*>(state->geom);
	char* image = static_cast<char*>(state->image);
	BlendFwdArgs ba{};
	ba.ranges = reinterpret_cast<const uint2*>(image + il.ranges);
	ba.point_list = reinterpret_cast<const uint32_t*>(state->binning);
	ba.records = reinterpret_cast<const float4*>(geom + gl.records);
	ba.W = W;
	ba.H = H;
	ba.grid_x = (W + TILE_X - 1) / TILE_X;
	ba.grid_y = (H + TILE_Y - 1) / TILE_Y;
	ba.n_contrib = reinterpret_cast<uint32_t*>(image + il.n_contrib);
	BRS_CUDA(launch_count_pairs(ba, out, stream));
	return BRS_OK;
}

int brs_sort_pairs_u32(const uint32_t* keys_in, const uint32_t* vals_in, uint32_t* keys_out, uint32_t* vals_out, int n,
                       int begin_bit, int end_bit, void* scratch, brs_stream stream)
{
	if (n < 0 || begin_bit < 0 || end_bit > 32 || begin_bit > end_bit)
		return BRS_ERR_INVALID_ARG;
	if (n == 0)
		return BRS_OK;
	if (n > (1 << 30))
		return BRS_ERR_UNSUPPORTED;
	if (keys_in == nullptr || keys_out == nullptr || vals_out == nullptr || scratch == nullptr)
		return BRS_ERR_INVALID_ARG;
	BRS_CUDA(sort_pairs(keys_in, vals_in, keys_out, vals_out, (size_t)n, begin_bit, end_bit, scratch, stream));
	return BRS_OK;
}

int brs_forward(const brs_view* view, const brs_gaussians* g, float* out_color, float* out_depth, int* radii,
                brs_alloc_fn alloc, void* alloc_ctx, brs_fwd_state* state, brs_stream stream)
{
	int st = validate_view(view, true);
	if (st != BRS_OK)
		return st;
	st = validate_gaussians(view, g);
	if (st != BRS_OK)
		return st;
	if (alloc == nullptr || state == nullptr)
		return BRS_ERR_INVALID_ARG;
	const int W = view->image_width, H = view->image_height, P = g->P;
	const size_t npix = (size_t)W * H;
	if ((npix > 0 && (out_color == nullptr || out_depth == nullptr)) || (P > 0 && radii == nullptr))
		return BRS_ERR_INVALID_ARG;
	const bool debug = view->debug != 0;

	memset(state, 0, sizeof(*state));

	if (P == 0) {
		// reference rasterize_points.cu:68-82: outputs are zero-filled and no kernel runs (so the
		// background is NOT composited when there are no Gaussians at all).
		if (npix > 0) {
			BRS_CUDA(cudaMemsetAsync(out_color, 0, sizeof(float) * NUM_CHANNELS * npix, stream));
			BRS_CUDA(cudaMemsetAsync(out_depth, 0, sizeof(float) * npix, stream));
		}
		return BRS_OK;
	}

	const uint32_t grid_x = (W + TILE_X - 1) / TILE_X, grid_y = (H + TILE_Y - 1) / TILE_Y;
	const GeomLayout gl = geom_layout(P);
	const ImageLayout il = image_layout(W, H);

	char* geom = static_cast<char*>(alloc(alloc_ctx, BRS_BUF_GEOM, gl.total));
	char* image = static_cast<char*>(alloc(alloc_ctx, BRS_BUF_IMAGE, il.total));
	if (geom == nullptr || image == nullptr)
		return BRS_ERR_ALLOC;
	state->geom = geom;
	state->geom_bytes = gl.total;
	state->image = image;
	state->image_bytes = il.total;

	uint32_t* d_total = reinterpret_cast<uint32_t*>(geom + gl.header);
	float4* records = reinterpret_cast<float4*>(geom + gl.records);
	uint32_t* depth_key = reinterpret_cast<uint32_t*>(geom + gl.depth_key);
	uint2* rect = reinterpret_cast<uint2*>(geom + gl.rect);
	uint32_t* order = reinterpret_cast<uint32_t*>(geom + gl.order);
	uint2* ranges = reinterpret_cast<uint2*>(image + il.ranges);
	float* final_T = reinterpret_cast<float*>(image + il.final_T);
	uint32_t* n_contrib = reinterpret_cast<uint32_t*>(image + il.n_contrib);

	BRS_CUDA(ensure_slot());
	BRS_CUDA(cudaMemsetAsync(geom + gl.header, 0, HEADER_BYTES, stream));

	PreprocessArgs pa{};
	pa.P = P;
	pa.D = view->sh_degree;
	pa.M = g->shs ? view->sh_coeffs : 0;
	pa.means3D = g->means3D;
	pa.scales = g->scales;
	pa.scale_modifier = view->scale_modifier;
	pa.rotations = g->rotations;
	pa.opacities = g->opacities;
	pa.shs = g->shs;
	pa.cov3D_precomp = g->cov3D_precomp;
	pa.colors_precomp = g->colors_precomp;
	pa.viewmatrix = view->viewmatrix;
	pa.projmatrix = view->projmatrix;
	pa.campos = view->campos;
	pa.W = W;
	pa.H = H;
	pa.tan_fovx = view->tanfovx;
	pa.tan_fovy = view->tanfovy;
	pa.focal_y = H / (2.0f * view->tanfovy); // rasterizer_impl.cu:223-224
	pa.focal_x = W / (2.0f * view->tanfovx);
	pa.grid_x = grid_x;
	pa.grid_y = grid_y;
	pa.prefiltered = view->prefiltered;
	pa.radii = radii;
	pa.records = records;
	pa.depth_key = depth_key;
	pa.rect = rect;
	pa.total_tiles = d_total;
	BRS_STAGE(BRS_STAGE_PREPROCESS, launch_preprocess(pa, stream), debug, stream);

	// R, R1 and the depth-key range leave for the host now.  The first radix pass of the depth sort (low
	// 8 key bits) needs none of them and keeps the GPU busy during the host round trip.
	BRS_CUDA(cudaMemcpyAsync(t_slot.pinned, d_total, 5 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
	BRS_CUDA(cudaEventRecord(t_slot.event, stream));

	char* scratch1 = static_cast<char*>(alloc(alloc_ctx, BRS_BUF_SCRATCH, depth_scratch_bytes(P)));
	if (scratch1 == nullptr)
		return BRS_ERR_ALLOC;
	const size_t pb = align_up(sizeof(uint32_t) * (size_t)P, 256);
	uint32_t* sorted_depth = reinterpret_cast<uint32_t*>(scratch1);
	uint32_t* first_keys = reinterpret_cast<uint32_t*>(scratch1 + pb);
	uint32_t* first_vals = reinterpret_cast<uint32_t*>(scratch1 + 2 * pb);
	char* depth_sort_scratch = scratch1 + 3 * pb;
	// culled Gaussians (key 0xFFFFFFFF) are dropped here: the later passes and the emission only see
	// the V visible ones
	BRS_STAGE(BRS_STAGE_DEPTH_SORT,
	          sort_pass(depth_key, nullptr, first_keys, first_vals, (size_t)P, 0u, 0, 8, depth_sort_scratch, stream, true,
	                    DEPTH_KEY_CULLED),
	          debug, stream);

	BRS_CUDA(cudaEventSynchronize(t_slot.event)); // the one host wait (reference: rasterizer_impl.cu:282)
	const uint32_t R = t_slot.pinned[0], R1 = t_slot.pinned[1];
	const uint32_t key_min = ~t_slot.pinned[2], key_max = t_slot.pinned[3];
	const uint32_t V = t_slot.pinned[4]; // visible Gaussians = entries that survived the first pass
	if (R > (1u << 30))
		return BRS_ERR_UNSUPPORTED;
	state->num_rendered = (int)R;

	// Remaining passes of the depth sort.  Visible keys lie in [key_min, key_max]; subtracting a bias
	// that is a multiple of 256 keeps the first pass's digit, preserves order and ties, and leaves only
	// bit_length(key_max - bias) significant bits (23-24 for a scene a few units deep instead of 32).
	// `order` ends up holding the V visible ids.  A visible Gaussian cannot carry the culled marker as
	// its depth bits: 0xFFFFFFFF is a NaN pattern the GPU's arithmetic never produces (its NaN is
	// 0x7FFFFFFF).
	if (V > 0) {
		const uint32_t bias = key_min <= key_max ? (key_min & ~0xFFu) : 0u;
		const uint32_t span = key_min <= key_max ? key_max - bias : 0u;
		int nbits = 8;
		while (nbits < 32 && (span >> nbits) != 0u)
			nbits++;
		const int rest = nbits > 8 ? nbits - 8 : 1;
		const int passes = (rest + 7) / 8;
		const int base_bits = rest / passes, extra = rest % passes;
		uint32_t *tmp_keys = nullptr, *tmp_vals = nullptr;
		sort_tmp_buffers(depth_sort_scratch, (size_t)P, &tmp_keys, &tmp_vals);
		const uint32_t* kin = first_keys;
		const uint32_t* vin = first_vals;
		int shift = 8;
		for (int p = 0; p < passes; p++) {
			const int bits = base_bits + (p < extra ? 1 : 0);
			const bool to_out = ((passes - 1 - p) & 1) == 0;
			uint32_t* ko = to_out ? sorted_depth : tmp_keys;
			uint32_t* vo = to_out ? order : tmp_vals;
			BRS_STAGE(BRS_STAGE_DEPTH_SORT, sort_pass(kin, vin, ko, vo, (size_t)V, bias, shift, bits, depth_sort_scratch, stream),
			          debug, stream);
			kin = ko;
			vin = vo;
			shift += bits;
		}
	}

	char* binning = static_cast<char*>(alloc(alloc_ctx, BRS_BUF_BINNING, binning_bytes(R)));
	if (binning == nullptr)
		return BRS_ERR_ALLOC;
	state->binning = binning;
	state->binning_bytes = binning_bytes(R);
	uint32_t* point_list = reinterpret_cast<uint32_t*>(binning);

	if (grid_x * grid_y > 0) {
		char* scratch2 = static_cast<char*>(alloc(alloc_ctx, BRS_BUF_SCRATCH, instance_scratch_bytes(P, R1, grid_x, grid_y)));
		if (scratch2 == nullptr)
			return BRS_ERR_ALLOC;
		const size_t rb = align_up(sizeof(uint32_t) * (size_t)R1, 256);
		uint32_t* cell_keys = reinterpret_cast<uint32_t*>(scratch2);
		uint32_t* cell_ids = reinterpret_cast<uint32_t*>(scratch2 + rb);
		uint32_t* sorted_keys = reinterpret_cast<uint32_t*>(scratch2 + 2 * rb);
		uint32_t* coarse_list = reinterpret_cast<uint32_t*>(scratch2 + 3 * rb);
		char* sort_scratch = scratch2 + 4 * rb;
		char* emit_scratch = sort_scratch + sort_scratch_bytes(R1);
		char* fine_scratch = emit_scratch + emit_scratch_bytes(P);
		const uint32_t ns_x = supertiles(grid_x), ns = ns_x * supertiles(grid_y);

		if (R1 > 0) {
			BRS_STAGE(BRS_STAGE_COARSE_EMIT,
			          launch_emit(order, rect, (size_t)V, ST_SHIFT, ns_x, cell_keys, cell_ids, (size_t)R1, emit_scratch,
			                      stream),
			          debug, stream);
			BRS_STAGE(BRS_STAGE_COARSE_SORT,
			          sort_pairs(cell_keys, cell_ids, sorted_keys, coarse_list, (size_t)R1, 0, tile_bits(ns), sort_scratch,
			                     stream),
			          debug, stream);
		}
		// also writes the (0,0) ranges of empty tiles (reference: cudaMemset, rasterizer_impl.cu:311)
		BRS_STAGE(BRS_STAGE_FINE_BIN,
		          launch_fine_binning(sorted_keys, coarse_list, (size_t)R1, rect, grid_x, grid_y, point_list, ranges,
		                              fine_scratch, stream),
		          debug, stream);
	}

	BlendFwdArgs ba{};
	ba.ranges = ranges;
	ba.point_list = point_list;
	ba.records = records;
	ba.bg = view->bg;
	ba.W = W;
	ba.H = H;
	ba.grid_x = grid_x;
	ba.grid_y = grid_y;
	ba.final_T = final_T;
	ba.n_contrib = n_contrib;
	ba.out_color = out_color;
	ba.out_depth = out_depth;
	Companion comp;
	if (!debug && companion_for(stream, comp)) {
		BRS_CUDA(cudaEventRecord(comp.before, stream));
		BRS_CUDA(cudaStreamWaitEvent(comp.stream, comp.before, 0));
		BRS_CUDA(launch_blend_forward(ba, comp.stream));
		BRS_CUDA(cudaEventRecord(comp.after, comp.stream));
		BRS_CUDA(cudaStreamWaitEvent(stream, comp.after, 0));
	} else {
		BRS_STAGE(BRS_STAGE_BLEND_FWD, launch_blend_forward(ba, stream), debug, stream);
	}
	return BRS_OK;
}

int brs_backward(const brs_view* view, const brs_gaussians* g, const int* radii, const brs_fwd_state* state,
                 const float* dL_dout_color, const float* dL_dout_depth, const brs_grads* grads, brs_alloc_fn alloc,
                 void* alloc_ctx, brs_stream stream)
{
	// reference: dL_dout_depth is plumbed and never used (backward.cu:443-554); it is read only when
	// the caller opts in with brs_grads.depth_gradient
	int st = validate_view(view, true);
	if (st != BRS_OK)
		return st;
	st = validate_gaussians(view, g);
	if (st != BRS_OK)
		return st;
	if (grads == nullptr || state == nullptr || alloc == nullptr)
		return BRS_ERR_INVALID_ARG;
	const int P = g->P;
	if (P == 0)
		return BRS_OK;
	const int W = view->image_width, H = view->image_height;
	const int M = g->shs ? view->sh_coeffs : 0;
	const bool acc = grads->accumulate != 0;
	const bool has_sr = g->scales != nullptr;
	// plain mode needs every tensor; in accumulate mode every parameter sink is optional (a NULL sink
	// = a frozen parameter whose gradient the caller does not want) and only dL_dmeans2D is required
	if (radii == nullptr || grads->dL_dmeans2D == nullptr)
		return BRS_ERR_INVALID_ARG;
	if (!acc && (grads->dL_dopacity == nullptr || grads->dL_dmeans3D == nullptr || (M > 0 && grads->dL_dsh == nullptr) ||
	             grads->dL_dcolors == nullptr || grads->dL_dcov3D == nullptr || grads->dL_dscales == nullptr ||
	             grads->dL_drotations == nullptr))
		return BRS_ERR_INVALID_ARG;
	const GeomLayout gl = geom_layout(P);
	const ImageLayout il = image_layout(W, H);
	const int R = state->num_rendered;
	if (state->geom == nullptr || state->image == nullptr || state->geom_bytes < gl.total ||
	    state->image_bytes < il.total || R < 0 || (R > 0 && (state->binning == nullptr || state->binning_bytes < binning_bytes(R))))
		return BRS_ERR_STATE;
	if ((size_t)W * H > 0 && dL_dout_color == nullptr)
		return BRS_ERR_INVALID_ARG;
	const bool depth_grad = grads->depth_gradient != 0;
	if (depth_grad && (size_t)W * H > 0 && (dL_dout_depth == nullptr || grads->out_depth == nullptr))
		return BRS_ERR_INVALID_ARG;
	const bool debug = view->debug != 0;
	const uint32_t grid_x = (W + TILE_X - 1) / TILE_X, grid_y = (H + TILE_Y - 1) / TILE_Y;

	const char* geom = static_cast<const char*>(state->geom);
	const char* image = static_cast<const char*>(state->image);

	float* accum = static_cast<float*>(alloc(alloc_ctx, BRS_BUF_SCRATCH, brs_backward_scratch_bytes(P)));
	if (accum == nullptr)
		return BRS_ERR_ALLOC;
	BRS_CUDA(cudaMemsetAsync(accum, 0, sizeof(float) * ACCUM_STRIDE * (size_t)P, stream));

	if (R > 0 && grid_x * grid_y > 0) {
		BlendBwdArgs bb{};
		bb.ranges = reinterpret_cast<const uint2*>(image + il.ranges);
		bb.point_list = reinterpret_cast<const uint32_t*>(state->binning);
		bb.records = reinterpret_cast<const float4*>(geom + gl.records);
		bb.bg = view->bg;
		bb.W = W;
		bb.H = H;
		bb.grid_x = grid_x;
		bb.grid_y = grid_y;
		bb.final_T = reinterpret_cast<const float*>(image + il.final_T);
		bb.n_contrib = reinterpret_cast<const uint32_t*>(image + il.n_contrib);
		bb.dL_dpixels = dL_dout_color;
		bb.dL_ddepth = depth_grad ? dL_dout_depth : nullptr;
		bb.out_depth = depth_grad ? grads->out_depth : nullptr;
		bb.accum = accum;
		Companion comp;
		if (!debug && companion_for(stream, comp)) {
			BRS_CUDA(cudaEventRecord(comp.before, stream));
			BRS_CUDA(cudaStreamWaitEvent(comp.stream, comp.before, 0));
			BRS_CUDA(launch_blend_backward(bb, comp.stream));
			BRS_CUDA(cudaEventRecord(comp.after, comp.stream));
			BRS_CUDA(cudaStreamWaitEvent(stream, comp.after, 0));
		} else {
			BRS_STAGE(BRS_STAGE_BLEND_BWD, launch_blend_backward(bb, stream), debug, stream);
		}
	}

	PreprocessBwdArgs pb{};
	pb.P = P;
	pb.D = view->sh_degree;
	pb.M = M;
	pb.means3D = g->means3D;
	pb.radii = radii;
	pb.shs = g->shs;
	pb.scales = g->scales;
	pb.rotations = g->rotations;
	pb.scale_modifier = view->scale_modifier;
	pb.cov3D_precomp = g->cov3D_precomp;
	pb.viewmatrix = view->viewmatrix;
	pb.projmatrix = view->projmatrix;
	pb.campos = view->campos;
	pb.W = W;
	pb.H = H;
	pb.tan_fovx = view->tanfovx;
	pb.tan_fovy = view->tanfovy;
	pb.focal_y = H / (2.0f * view->tanfovy);
	pb.focal_x = W / (2.0f * view->tanfovx);
	pb.accum = accum;
	pb.accumulate = acc ? 1 : 0;
	pb.depth_gradient = depth_grad ? 1 : 0;
	pb.dL_dmeans2D = grads->dL_dmeans2D;
	pb.dL_dcolors = grads->dL_dcolors;
	pb.dL_dopacity = grads->dL_dopacity;
	pb.dL_dmeans3D = grads->dL_dmeans3D;
	pb.dL_dcov3D = grads->dL_dcov3D;
	pb.dL_dsh = M > 0 ? grads->dL_dsh : nullptr;
	pb.dL_dscales = grads->dL_dscales;
	pb.dL_drotations = grads->dL_drotations;
	// (measured: moving this kernel to the companion stream as well is slightly slower)
	BRS_STAGE(BRS_STAGE_PREPROCESS_BWD, launch_preprocess_backward(pb, stream), debug, stream);
	return BRS_OK;
}

int brs_visible_filter(const brs_view* view, int P, const float* means3D, const float* scales, int scales_stride,
                       const float* rotations, const float* cov3D_precomp, int* radii, brs_stream stream)
{
	int st = validate_view(view, false);
	if (st != BRS_OK)
		return st;
	if (P < 0)
		return BRS_ERR_INVALID_ARG;
	if (P == 0)
		return BRS_OK;
	const bool has_sr = scales != nullptr && rotations != nullptr;
	if (means3D == nullptr || radii == nullptr || has_sr == (cov3D_precomp != nullptr) || (has_sr && scales_stride < 3))
		return BRS_ERR_INVALID_ARG;
	const int W = view->image_width, H = view->image_height;
	FilterArgs fa{};
	fa.P = P;
	fa.means3D = means3D;
	fa.scales = scales;
	fa.scales_stride = scales_stride;
	fa.scale_modifier = view->scale_modifier;
	fa.rotations = rotations;
	fa.cov3D_precomp = cov3D_precomp;
	fa.viewmatrix = view->viewmatrix;
	fa.projmatrix = view->projmatrix;
	fa.W = W;
	fa.H = H;
	fa.tan_fovx = view->tanfovx;
	fa.tan_fovy = view->tanfovy;
	fa.focal_y = H / (2.0f * view->tanfovy);
	fa.focal_x = W / (2.0f * view->tanfovx);
	fa.grid_x = (W + TILE_X - 1) / TILE_X;
	fa.grid_y = (H + TILE_Y - 1) / TILE_Y;
	fa.prefiltered = view->prefiltered;
	fa.radii = radii;
	BRS_STAGE(BRS_STAGE_PREPROCESS, launch_filter(fa, stream), view->debug != 0, stream);
	return BRS_OK;
}

int brs_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, uint8_t* present,
                     brs_stream stream)
{
	(void)projmatrix; // reference in_frustum computes p_proj but only tests view-space z (auxiliary.h:154)
	if (P < 0)
		return BRS_ERR_INVALID_ARG;
	if (P == 0)
		return BRS_OK;
	if (means3D == nullptr || viewmatrix == nullptr || present == nullptr)
		return BRS_ERR_INVALID_ARG;
	BRS_CUDA(launch_check_frustum(P, means3D, viewmatrix, present, stream));
	return BRS_OK;
}

} // extern "C"
