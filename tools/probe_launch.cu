// What does one kernel boundary cost on this GPU?  Chains of N dependent kernels on one stream, each doing
// one global round trip (load -> store) with a small grid, timed with CUDA events:
//   plain   : ordinary launches
//   pdl     : programmatic dependent launch (griddepcontrol.wait at the top, launch_dependents at once)
//   graph   : the plain chain captured in a CUDA graph
//   graphpdl: the PDL chain captured in a CUDA graph
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/probe_launch.cu -o bloomscene_b200/_build/probe_launch
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                                          \
	do {                                                                                                               \
		cudaError_t e_ = (x);                                                                                          \
		if (e_ != cudaSuccess) {                                                                                       \
			fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                                 \
			exit(1);                                                                                                   \
		}                                                                                                              \
	} while (0)

template <bool PDL>
__global__ void __launch_bounds__(256) hop_kernel(const unsigned* in, unsigned* out, int n)
{
	if (PDL) {
		asm volatile("griddepcontrol.wait;" ::: "memory");
		asm volatile("griddepcontrol.launch_dependents;");
	}
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n)
		out[i] = in[i] + 1u;
}

template <bool PDL>
static void launch_chain(unsigned* a, unsigned* b, int n, int len, int blocks, cudaStream_t s)
{
	for (int k = 0; k < len; k++) {
		cudaLaunchConfig_t cfg = {};
		cfg.gridDim = dim3(blocks);
		cfg.blockDim = dim3(256);
		cfg.stream = s;
		cudaLaunchAttribute attr;
		attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
		attr.val.programmaticStreamSerializationAllowed = 1;
		cfg.attrs = &attr;
		cfg.numAttrs = PDL ? 1 : 0;
		const unsigned* in = (k & 1) ? b : a;
		unsigned* out = (k & 1) ? a : b;
		CK(cudaLaunchKernelEx(&cfg, hop_kernel<PDL>, in, out, n));
	}
}

static float median(std::vector<float> v)
{
	std::sort(v.begin(), v.end());
	return v[v.size() / 2];
}

int main()
{
	cudaStream_t s;
	CK(cudaStreamCreate(&s));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	const int len = 12;
	for (int blocks : {32, 296, 1184}) {
		const int n = blocks * 256;
		unsigned *a, *b;
		CK(cudaMalloc(&a, 4ull * n));
		CK(cudaMalloc(&b, 4ull * n));
		CK(cudaMemset(a, 0, 4ull * n));
		cudaGraphExec_t gexec[2];
		for (int pdl = 0; pdl < 2; pdl++) {
			cudaGraph_t g;
			CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
			if (pdl)
				launch_chain<true>(a, b, n, len, blocks, s);
			else
				launch_chain<false>(a, b, n, len, blocks, s);
			CK(cudaStreamEndCapture(s, &g));
			CK(cudaGraphInstantiate(&gexec[pdl], g, 0));
		}
		std::vector<float> t[4];
		for (int it = 0; it < 60; it++) {
			for (int mode = 0; mode < 4; mode++) {
				CK(cudaStreamSynchronize(s));
				CK(cudaEventRecord(e0, s));
				if (mode == 0)
					launch_chain<false>(a, b, n, len, blocks, s);
				else if (mode == 1)
					launch_chain<true>(a, b, n, len, blocks, s);
				else
					CK(cudaGraphLaunch(gexec[mode - 2], s));
				CK(cudaEventRecord(e1, s));
				CK(cudaEventSynchronize(e1));
				float ms;
				CK(cudaEventElapsedTime(&ms, e0, e1));
				if (it >= 10)
					t[mode].push_back(ms);
			}
		}
		printf("{\"chain\": %d, \"blocks\": %d, \"us_per_kernel\": {\"plain\": %.2f, \"pdl\": %.2f, \"graph\": %.2f, \"graph_pdl\": %.2f}}\n", len,
		       blocks, 1e3 * median(t[0]) / len, 1e3 * median(t[1]) / len, 1e3 * median(t[2]) / len, 1e3 * median(t[3]) / len);
		cudaFree(a);
		cudaFree(b);
	}
	return 0;
}
