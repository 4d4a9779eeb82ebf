// Diagnostic: fe_sqr vs fe_mul(a,a) when some limbs are compile-time constants.
#include <cstdio>
#include "../../bulletproof-perm_b200/csrc/ge25519.cuh"
__device__ void dbl_ref(ge_ext &r, const ge_ext &p) {
    fe a, b, c, e, f, g, h, t;
    fe_mul(a, p.X, p.X); fe_mul(b, p.Y, p.Y); fe_mul(c, p.Z, p.Z); fe_dbl(c, c);
    fe_add(h, a, b); fe_add(t, p.X, p.Y); fe_mul(t, t, t); fe_sub(e, h, t); fe_sub(g, a, b); fe_add(f, c, g);
    fe_mul(r.X, e, f); fe_mul(r.Y, g, h); fe_mul(r.Z, f, g); fe_mul(r.T, e, h);
}
__global__ void k(const uint32_t *in, uint32_t *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y;
    for (int k = 0; k < 8; k++) { x.v[k] = in[16 * i + k]; y.v[k] = in[16 * i + 8 + k]; }
    uint32_t bad = 0;
    fe r1, r2, c1, c2;
    // 1: (x0,0,...,0)
    fe z; fe_set0(z); z.v[0] = x.v[0];
    fe_sqr(r1, z); fe_mul(r2, z, z); fe_canon(c1, r1); fe_canon(c2, r2);
    for (int k = 0; k < 8; k++) if (c1.v[k] != c2.v[k]) bad |= 1;
    // 2: const 1
    fe_set1(z);
    fe_sqr(r1, z); fe_mul(r2, z, z); fe_canon(c1, r1); fe_canon(c2, r2);
    for (int k = 0; k < 8; k++) if (c1.v[k] != c2.v[k]) bad |= 2;
    // 3: full runtime
    fe_sqr(r1, x); fe_mul(r2, x, x); fe_canon(c1, r1); fe_canon(c2, r2);
    for (int k = 0; k < 8; k++) if (c1.v[k] != c2.v[k]) bad |= 4;
    // 4: doubling with Z = 1
    ge_ext P, A, B;
    P.X = x; P.Y = y; fe_set1(P.Z); fe_mul(P.T, x, y);
    ge_double(A, P); dbl_ref(B, P);
    const fe *pa = &A.X, *pb = &B.X;
    for (int j = 0; j < 4; j++) { fe_canon(c1, pa[j]); fe_canon(c2, pb[j]); for (int k = 0; k < 8; k++) if (c1.v[k] != c2.v[k]) bad |= (16u << j); }
    // 5: doubling, all runtime
    P.Z = y;
    ge_double(A, P); dbl_ref(B, P);
    for (int j = 0; j < 4; j++) { fe_canon(c1, pa[j]); fe_canon(c2, pb[j]); for (int k = 0; k < 8; k++) if (c1.v[k] != c2.v[k]) bad |= (256u << j); }
    // 6: p2 doubling
    ge_double_p2(A, P);
    for (int j = 0; j < 3; j++) { fe_canon(c1, pa[j]); fe_canon(c2, pb[j]); for (int k = 0; k < 8; k++) if (c1.v[k] != c2.v[k]) bad |= (4096u << j); }
    out[i] = bad;
}
int main() {
    const int n = 4096;
    uint32_t *h = new uint32_t[16 * n], *d, *o, *ho = new uint32_t[n];
    srand(7);
    for (int i = 0; i < 16 * n; i++) h[i] = ((uint32_t)rand() << 16) ^ (uint32_t)rand();
    cudaMalloc(&d, 64 * n); cudaMalloc(&o, 4 * n);
    cudaMemcpy(d, h, 64 * n, cudaMemcpyHostToDevice);
    k<<<n / 128, 128>>>(d, o, n);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(ho, o, 4 * n, cudaMemcpyDeviceToHost);
    uint32_t all = 0; int cnt = 0;
    for (int i = 0; i < n; i++) { all |= ho[i]; cnt += ho[i] != 0; }
    printf("err=%s badmask=0x%x count=%d first=0x%x\n", cudaGetErrorString(e), all, cnt, ho[0]);
}
