// Measurement only: cost of a dependent kernel boundary inside a CUDA graph on this GPU, with and without programmatic
// dependent launch (PDL: cudaLaunchAttributeProgrammaticStreamSerialization + griddepcontrol.wait at kernel entry).
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tools/_build/pdl_probe tools/pdl_probe.cu && tools/_build/pdl_probe
#include <cstdio>
#include <cuda_runtime.h>

template <bool PDL>
__global__ void __launch_bounds__(256) stream_kernel(const float4* __restrict__ in, float4* __restrict__ out, long long n4, float k) {
    if (PDL) {
        asm volatile("griddepcontrol.launch_dependents;");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 v = in[i];
        v.x *= k; v.y *= k; v.z *= k; v.w *= k;
        out[i] = v;
    }
}

template <bool PDL>
static void launch(cudaStream_t st, int grid, const float4* in, float4* out, long long n4) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    cudaLaunchKernelEx(&cfg, stream_kernel<PDL>, in, out, n4, 1.0001f);
}

template <bool PDL>
static float run(int grid, long long n4, int chain, float4* a, float4* b) {
    cudaStream_t st; cudaStreamCreate(&st);
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < chain; ++i) launch<PDL>(st, grid, (i & 1) ? b : a, (i & 1) ? a : b, n4);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(e0, st);
    for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(e1, st);
    cudaStreamSynchronize(st);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error: %s\n", cudaGetErrorString(err));
    cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    return ms * 1000.f / (5.f * chain);
}

int main() {
    const long long maxn4 = (64ll << 20) / 16;
    float4 *a, *b;
    cudaMalloc(&a, maxn4 * 16); cudaMalloc(&b, maxn4 * 16);
    cudaMemset(a, 0, maxn4 * 16); cudaMemset(b, 0, maxn4 * 16);
    struct { const char* name; int grid; long long bytes; } cases[] = {
        {"tiny   (1 CTA, 4 KB)", 1, 4096}, {"small  (37 CTAs, 1 MB)", 37, 1 << 20}, {"medium (592 CTAs, 16 MB)", 592, 16 << 20},
        {"large  (592 CTAs, 64 MB)", 592, 64 << 20}};
    for (auto& c : cases) {
        const float t0 = run<false>(c.grid, c.bytes / 16, 500, a, b);
        const float t1 = run<true>(c.grid, c.bytes / 16, 500, a, b);
        printf("%-28s  plain %7.2f us/kernel   PDL %7.2f us/kernel   saved %5.2f us\n", c.name, t0, t1, t0 - t1);
    }
    return 0;
}
