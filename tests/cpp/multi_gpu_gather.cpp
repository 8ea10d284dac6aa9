// A C++ host that shards a batch of tracking_step problems over the GPUs of one box WITHOUT Python: one context per device,
// the device-pointer entry point on every rank's stream, the kernel-packed records gathered by tdlo_all_gather_packed over an
// NCCL communicator the host created (single process, ncclCommInitAll).  Driven by tests/test_multi_gpu_c.py, which writes the
// inputs and compares the gathered records with a single-GPU run.   usage: multi_gpu_gather <in.bin> <out.bin>   exit 77: < 2 GPUs
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "trackdlo_b200.h"

#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)
#define CKN(x) do { ncclResult_t e_ = (x); if (e_ != ncclSuccess) { fprintf(stderr, "%s: %s\n", #x, ncclGetErrorString(e_)); return 3; } } while (0)
#define CKT(c, x) do { int e_ = (x); if (e_ != 0) { fprintf(stderr, "%s: %d %s\n", #x, e_, tdlo_last_error(c)); return 4; } } while (0)

template <class T> static std::vector<T> rd(FILE* f, size_t n) { std::vector<T> v(n); if (n && fread(v.data(), sizeof(T), n, f) != n) { fprintf(stderr, "short read\n"); exit(5); } return v; }
template <class T> static T* up(const T* h, size_t n) { T* d = nullptr; cudaMalloc(&d, (n ? n : 1) * sizeof(T)); if (n) cudaMemcpy(d, h, n * sizeof(T), cudaMemcpyHostToDevice); return d; }

int main(int argc, char** argv) {
    if (argc < 3) return 1;
    int ngpu = 0;
    CKC(cudaGetDeviceCount(&ngpu));
    if (ngpu < 2) return 77;
    const int world = 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 1;
    const auto hdr = rd<int64_t>(f, 2);
    const int F = (int)hdr[0], N = (int)hdr[1];
    const auto xo = rd<int64_t>(f, F + 1);
    const auto X = rd<double>(f, (size_t)xo[F] * 3);
    const auto Y = rd<double>(f, (size_t)F * N * 3);
    const auto rest = rd<double>(f, (size_t)F * N);
    const auto vo = rd<int64_t>(f, F + 1);
    const auto vis = rd<int32_t>(f, (size_t)vo[F]);
    const auto eo = rd<int64_t>(f, F + 1);
    const auto ext = rd<int32_t>(f, (size_t)eo[F]);
    const auto par = rd<tdlo_track_params>(f, 1);
    fclose(f);
    const int Fr = F / world;                       // frames per rank (the test sends a multiple of the world size)
    const size_t rec = 3 * (size_t)N + 4;

    ncclComm_t comms[2];
    int devs[2] = {0, 1};
    CKN(ncclCommInitAll(comms, world, devs));
    tdlo_ctx* ctx[2]; cudaStream_t st[2]; double* packed[2]; double* dY[2];
    for (int r = 0; r < world; r++) {
        CKC(cudaSetDevice(r));
        CKC(cudaStreamCreate(&st[r]));
        const int f0 = r * Fr;
        std::vector<int64_t> x1(Fr + 1), v1(Fr + 1), e1(Fr + 1);
        for (int i = 0; i <= Fr; i++) { x1[i] = xo[f0 + i] - xo[f0]; v1[i] = vo[f0 + i] - vo[f0]; e1[i] = eo[f0 + i] - eo[f0]; }
        CKT(nullptr, tdlo_create(&ctx[r], r, Fr, N, x1[Fr]));
        tdlo_track_batch b{};
        b.n_frames = Fr; b.n_nodes = N;
        b.X = up(X.data() + xo[f0] * 3, (size_t)x1[Fr] * 3); b.x_offsets = up(x1.data(), x1.size());
        dY[r] = up(Y.data() + (size_t)f0 * N * 3, (size_t)Fr * N * 3); b.Y = dY[r];
        std::vector<double> s2(Fr, 0.0);
        b.sigma2 = up(s2.data(), s2.size());
        b.geodesic_coord = up(rest.data() + (size_t)f0 * N, (size_t)Fr * N);
        b.visible = up(vis.data() + vo[f0], (size_t)v1[Fr]); b.visible_offsets = up(v1.data(), v1.size());
        b.visible_ext = up(ext.data() + eo[f0], (size_t)e1[Fr]); b.visible_ext_offsets = up(e1.data(), e1.size());
        CKC(cudaMalloc(&packed[r], (size_t)F * rec * sizeof(double)));
        CKC(cudaMemset(packed[r], 0, (size_t)F * rec * sizeof(double)));
        b.packed_results = packed[r] + (size_t)r * Fr * rec;           // this rank's slot of the gathered array
        CKT(ctx[r], tdlo_tracking_step_batched_device(ctx[r], &b, &par[0], st[r]));
    }
    CKN(ncclGroupStart());
    for (int r = 0; r < world; r++) { CKC(cudaSetDevice(r)); CKT(ctx[r], tdlo_all_gather_packed(ctx[r], comms[r], packed[r], Fr, N, r, st[r])); }
    CKN(ncclGroupEnd());
    std::vector<double> out0((size_t)F * rec), out1((size_t)F * rec);
    for (int r = 0; r < world; r++) {
        CKC(cudaSetDevice(r));
        CKC(cudaStreamSynchronize(st[r]));
        CKT(ctx[r], tdlo_synchronize(ctx[r]));
        CKC(cudaMemcpy(r == 0 ? out0.data() : out1.data(), packed[r], (size_t)F * rec * sizeof(double), cudaMemcpyDeviceToHost));
    }
    for (size_t i = 0; i < out0.size(); i++) if (out0[i] != out1[i]) { fprintf(stderr, "ranks disagree at %zu\n", i); return 6; }
    FILE* g = fopen(argv[2], "wb");
    fwrite(out0.data(), sizeof(double), out0.size(), g);
    fclose(g);
    for (int r = 0; r < world; r++) { cudaSetDevice(r); tdlo_destroy(ctx[r]); ncclCommDestroy(comms[r]); }
    printf("gathered %d frames x %zu doubles on %d GPUs\n", F, rec, world);
    return 0;
}
