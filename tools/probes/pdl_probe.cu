// Kernel-boundary cost probe (B200): a chain of persistent-style kernels, one CTA per SM, each spinning `spin_ns` and then writing
// 64 KiB; consecutive launches are linked by programmatic dependent launch.  Two shared-memory footprints: 200 KiB (the next
// kernel's CTAs cannot become resident before this kernel's CTAs exit) and 100 KiB (they can: prologue overlaps).  Reports, per
// boundary, when the dependent's CTAs entered and when their griddepcontrol.wait returned relative to the last exit of the primary.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probes/pdl_probe tools/probes/pdl_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// stamps[launch][cta][3] = entry, after wait, exit
__global__ void __launch_bounds__(256, 1) chain_kernel(unsigned long long* stamps, float* sink, int launch, int spin_ns, int pdl, int prologue_ns) {
    extern __shared__ unsigned char smem[];
    unsigned long long* my = stamps + (size_t(launch) * gridDim.x + blockIdx.x) * 3;
    const unsigned long long t_in = gtime();
    if (threadIdx.x == 0) my[0] = t_in;
    if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // stand-in for the prologue (barrier init, TMEM alloc, cluster sync, weight prefetch): independent of the previous grid
    while (gtime() - t_in < (unsigned long long)prologue_ns) {}
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    const unsigned long long t_go = gtime();
    if (threadIdx.x == 0) my[1] = t_go;
    // read something the previous launch wrote (keeps the dependency real), spin, write 64 KiB
    float acc = sink[(size_t((launch + 1) & 1) * gridDim.x + blockIdx.x) * 16384 + threadIdx.x];
    while (gtime() - t_go < (unsigned long long)spin_ns) {}
    float* out = sink + (size_t(launch & 1) * gridDim.x + blockIdx.x) * 16384;
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) out[i] = acc + float(i);
    __syncthreads();
    if (threadIdx.x == 0) my[2] = gtime();
    if (smem[threadIdx.x] == 77 && acc == 123.f) out[0] = 1.f;  // keep smem referenced
}

int main(int argc, char** argv) {
    const int launches = 24, grid = 148;
    unsigned long long* stamps;
    float* sink;
    cudaMalloc(&stamps, sizeof(unsigned long long) * launches * grid * 3);
    cudaMalloc(&sink, sizeof(float) * 2 * grid * 16384);
    cudaMemset(sink, 0, sizeof(float) * 2 * grid * 16384);
    cudaStream_t s;
    cudaStreamCreate(&s);
    cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    std::vector<unsigned long long> h(size_t(launches) * grid * 3);
    for (int spin_ns : {3000, 6000}) {
        for (int prologue_ns : {0, 1000}) {
            for (int mode = 0; mode < 5; ++mode) {
                // 0: no PDL, 200K; 1: PDL, 200K (exclusive); 2: PDL, 100K (co-resident); 3: PDL 100K grid 74 ; 4: no PDL 100K
                const int pdl = (mode == 0 || mode == 4) ? 0 : 1;
                const size_t smem = (mode == 0 || mode == 1) ? 200 * 1024 : 100 * 1024;
                const int g = mode == 3 ? 74 : grid;
                for (int rep = 0; rep < 2; ++rep) {
                    cudaMemsetAsync(stamps, 0, sizeof(unsigned long long) * launches * grid * 3, s);
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0);
                    cudaEventCreate(&e1);
                    cudaEventRecord(e0, s);
                    for (int l = 0; l < launches; ++l) {
                        cudaLaunchConfig_t cfg{};
                        cfg.gridDim = dim3(g);
                        cfg.blockDim = dim3(256);
                        cfg.dynamicSmemBytes = smem;
                        cfg.stream = s;
                        cudaLaunchAttribute attr[1];
                        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
                        attr[0].val.programmaticStreamSerializationAllowed = 1;
                        cfg.attrs = attr;
                        cfg.numAttrs = pdl ? 1 : 0;
                        cudaLaunchKernelEx(&cfg, chain_kernel, stamps, sink, l, spin_ns, pdl, prologue_ns);
                    }
                    cudaEventRecord(e1, s);
                    cudaStreamSynchronize(s);
                    float ms = 0;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (rep == 0) continue;
                    cudaMemcpy(h.data(), stamps, sizeof(unsigned long long) * launches * grid * 3, cudaMemcpyDeviceToHost);
                    // per boundary l-1 -> l (skip the first 4): last exit of l-1, first/last entry of l, last go of l
                    double d_entry_first = 0, d_entry_last = 0, d_go_last = 0, period = 0;
                    int n = 0;
                    for (int l = 5; l < launches; ++l) {
                        unsigned long long last_exit = 0, first_entry = ~0ull, last_entry = 0, last_go = 0, prev_go = 0;
                        for (int c = 0; c < g; ++c) {
                            const unsigned long long* p = &h[(size_t(l - 1) * grid + c) * 3];
                            const unsigned long long* q = &h[(size_t(l) * grid + c) * 3];
                            last_exit = std::max(last_exit, p[2]);
                            prev_go = std::max(prev_go, p[1]);
                            first_entry = std::min(first_entry, q[0]);
                            last_entry = std::max(last_entry, q[0]);
                            last_go = std::max(last_go, q[1]);
                        }
                        d_entry_first += double((long long)(first_entry - last_exit));
                        d_entry_last += double((long long)(last_entry - last_exit));
                        d_go_last += double((long long)(last_go - last_exit));
                        period += double((long long)(last_go - prev_go));
                        ++n;
                    }
                    const char* names[5] = {"noPDL smem200K", "PDL   smem200K", "PDL   smem100K", "PDL   smem100K grid74", "noPDL smem100K"};
                    printf("spin %d ns prologue %d ns  %-22s: period %.2f us (event %.2f us/launch) | vs last exit of the previous grid: first entry %+.2f us, last entry %+.2f us, last go %+.2f us\n",
                           spin_ns, prologue_ns, names[mode], period / n / 1e3, ms * 1e3 / launches, d_entry_first / n / 1e3, d_entry_last / n / 1e3, d_go_last / n / 1e3);
                    fflush(stdout);
                }
            }
        }
    }
    return 0;
}
