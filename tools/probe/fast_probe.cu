#include "../../airdos_b200/csrc/orb.cu"
// score one 44x40 tile with the device functions and with a plain CPU loop
__global__ void k(const uint8_t* img, int w, int h, int t, int* out) {
    int x = threadIdx.x + 3, y = blockIdx.x + 3;
    if (x >= w - 3) return;
    const uint8_t* c = img + y * w + x; int bw = w; int v = c[0]; int d[16];
    d[0] = v - c[3 * bw];      d[1] = v - c[3 * bw + 1];  d[2] = v - c[2 * bw + 2];   d[3] = v - c[bw + 3];
    d[4] = v - c[3];           d[5] = v - c[-bw + 3];     d[6] = v - c[-2 * bw + 2];  d[7] = v - c[-3 * bw + 1];
    d[8] = v - c[-3 * bw];     d[9] = v - c[-3 * bw - 1]; d[10] = v - c[-2 * bw - 2]; d[11] = v - c[-bw - 3];
    d[12] = v - c[-3];         d[13] = v - c[bw - 3];     d[14] = v - c[2 * bw - 2];  d[15] = v - c[3 * bw - 1];
    uint32_t hi = 0, lo = 0;
    for (int i = 0; i < 16; ++i) { hi |= (uint32_t)(d[i] > t) << i; lo |= (uint32_t)(d[i] < -t) << i; }
    int s = 0;
    if (adb::has_run9(hi) || adb::has_run9(lo)) s = adb::fast_best(d) - 1;
    out[y * w + x] = s;
}
int main() {
    const int w = 44, h = 40, t = 7;
    std::vector<uint8_t> img(w * h);
    srand(1);
    for (auto& p : img) p = 100 + rand() % 60;
    uint8_t* d; int* o; cudaMalloc(&d, w * h); cudaMalloc(&o, w * h * 4); cudaMemset(o, 0, w * h * 4);
    cudaMemcpy(d, img.data(), w * h, cudaMemcpyHostToDevice);
    k<<<h - 6, 64>>>(d, w, h, t, o);
    std::vector<int> out(w * h); cudaMemcpy(out.data(), o, w * h * 4, cudaMemcpyDeviceToHost);
    const int dx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1}, dy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};
    int bad = 0;
    for (int y = 3; y < h - 3; ++y) for (int x = 3; x < w - 3; ++x) {
        int dd[16]; for (int i = 0; i < 16; ++i) dd[i] = img[y * w + x] - img[(y + dy[i]) * w + x + dx[i]];
        int best = -256;
        for (int s = 0; s < 16; ++s) { int lo = 999, hi = -999; for (int k2 = 0; k2 < 9; ++k2) { lo = std::min(lo, dd[(s + k2) & 15]); hi = std::max(hi, dd[(s + k2) & 15]); } best = std::max(best, std::max(lo, -hi)); }
        int ref = best > t ? best - 1 : 0;
        if (ref != out[y * w + x]) { if (bad < 5) printf("(%d,%d) ref %d got %d\n", x, y, ref, out[y * w + x]); ++bad; }
    }
    printf("fast_probe mismatches %d (%s)\n", bad, cudaGetErrorString(cudaGetLastError()));
}
