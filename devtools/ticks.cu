// development probe: per-phase latency of one warm fused evaluation (one warp alone on an SM)
#define NEO_TICKS
#include "../neo_planner_b200/csrc/minco_warp.cuh"
#include <cstdio>
#include <vector>
using namespace neo;
template <int M>
__global__ void k_ticks(DevParams P, MapView map, const double *x, const double *ht, long long *ticks, double *out)
{
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    WarpMem m = carve(smem, M);
    begin_problem(m, M, lane, ht, ht + 6);
    const int n = 3 * M - 2;
    double xl = lane < n ? x[lane] : 0.0;
    EvalOut ev;
    for (int rep = 0; rep < 3; rep++) eval_fg<SAMPLE_BY_PIECE>(P, map, m, M, lane, xl, true, ev, rep == 2 ? ticks : nullptr);
    if (lane == 0) { out[0] = ev.f; out[1] = ev.ns; }
    if (lane < n) out[2 + lane] = ev.g;
}
int main()
{
    const int H = 300, W = 300, M = 3;
    std::vector<Cell> cells(H * W);
    for (int r = 0; r < H; r++) for (int c = 0; c < W; c++) {
        double dx = (c - 150) * 0.1, dy = (r - 150) * 0.1, d = sqrt(dx * dx + dy * dy);
        cells[r * W + c] = Cell{d > 0 ? dx / d * 0.1 : 0, d > 0 ? dy / d * 0.1 : 0, d, 0};
    }
    Cell *dc; cudaMalloc(&dc, sizeof(Cell) * H * W); cudaMemcpy(dc, cells.data(), sizeof(Cell) * H * W, cudaMemcpyHostToDevice);
    MapView map{dc, H, W, 0.1, 0.0, -15.0, 10.0};
    DevParams P{1.0, 0.5, 5.0, 0.7, 0.1, 1, 1, 1, 10000, 5};
    // trajectory from (12,-1) to (17,1) passing near the obstacle centre (15,0); tau for T = 3.75, 2.5, 3.75
    double x[7] = {13.6, 15.3, -0.3, 0.4, 0.9555114450274365, -0.22314355131420976, 0.9555114450274365};
    double ht[12] = {12, -1, 0.5, 0.1, 0, 0, 17, 1, 0.8, 0.2, 0, 0};
    double *dx, *dht, *dout; long long *dt;
    cudaMalloc(&dx, sizeof(x)); cudaMalloc(&dht, sizeof(ht)); cudaMalloc(&dout, 64 * 8); cudaMalloc(&dt, 16 * 8);
    cudaMemcpy(dx, x, sizeof(x), cudaMemcpyHostToDevice); cudaMemcpy(dht, ht, sizeof(ht), cudaMemcpyHostToDevice);
    k_ticks<3><<<1, 32, sizeof(double) * warp_mem_doubles(M)>>>(P, map, dx, dht, dt, dout);
    long long t[16]; double out[16];
    cudaMemcpy(t, dt, sizeof(t), cudaMemcpyDeviceToHost); cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost);
    printf("err %s f=%g ns=%g\n", cudaGetErrorString(cudaGetLastError()), out[0], out[1]);
    const char *names[] = {"tau->T (exp)", "nodes: load+solve", "hermite coeffs", "energy", "sample loop + reduce", "h = H^T gC",
                           "adjoint elimination", "G rows", "grad_T, grad_tau"};
    for (int i = 0; i < 9; i++) printf("%-24s %7lld cycles\n", names[i], t[i + 1] - t[i]);
    printf("%-24s %7lld cycles\n", "TOTAL", t[9] - t[0]);

}
