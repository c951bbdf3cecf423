// Standalone validation + timing of the tcgen05 3xTF32 conv kernel against a double-accumulating reference kernel.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/test_conv_tc tools/test_conv_tc.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "legacy/conv_tc2.cuh"
#include "../tts_b200/csrc/conv_tc3.cuh"

using namespace b200tts::tc;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void ref_conv(const float* x, const float* w, const float* bias, const float* res, const float* yold, float* y,
                         int C, int Cout, int T, int K, int dil, int pad, float slope, float scale, int accum) {
    int t = blockIdx.x * blockDim.x + threadIdx.x, co = blockIdx.y, b = blockIdx.z;
    if (t >= T) return;
    double acc = 0;
    for (int ci = 0; ci < C; ++ci)
        for (int k = 0; k < K; ++k) {
            int ti = t + k * dil - pad;
            if (ti < 0 || ti >= T) continue;
            float v = x[((size_t)b * C + ci) * T + ti];
            v = v > 0 ? v : v * slope;
            acc += (double)v * (double)w[((size_t)co * C + ci) * K + k];
        }
    float u = (float)acc + bias[co];
    size_t o = ((size_t)b * Cout + co) * T + t;
    if (res) u += res[o];
    u *= scale;
    if (accum) u += yold[o];
    y[o] = u;
}

static std::vector<float> pack_tc(const std::vector<float>& W, int Cout, int Cin, int K, int N) {
    const int ntile = (Cout + N - 1) / N, nchunk = (Cin + KC - 1) / KC;
    const size_t blk = (size_t)2 * NSLAB * N * 4;  // floats per (tile, chunk, tap)
    std::vector<float> P((size_t)ntile * nchunk * K * blk, 0.f);
    for (int tile = 0; tile < ntile; ++tile)
        for (int c = 0; c < nchunk; ++c)
            for (int k = 0; k < K; ++k) {
                float* dst = P.data() + (((size_t)tile * nchunk + c) * K + k) * blk;
                for (int s = 0; s < NSLAB; ++s)
                    for (int n = 0; n < N; ++n)
                        for (int i = 0; i < 4; ++i) {
                            const int co = tile * N + n, ci = c * KC + 4 * s + i;
                            float v = (co < Cout && ci < Cin) ? W[((size_t)co * Cin + ci) * K + k] : 0.f;
                            uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u;
                            float hi; memcpy(&hi, &u, 4);
                            dst[((size_t)s * N + n) * 4 + i] = hi;
                            dst[((size_t)(NSLAB + s) * N + n) * 4 + i] = v - hi;
                        }
            }
    return P;
}

static std::vector<float> pack_grouped(const std::vector<float>& W, int Cout, int Cin, int K, int G) {
    const int J = (K + G - 1) / G, nchunk = (Cin + KC - 1) / KC;
    const size_t blk = (size_t)2 * NSLAB * 128 * 4;
    std::vector<float> P((size_t)nchunk * J * blk, 0.f);
    for (int c = 0; c < nchunk; ++c)
        for (int j = 0; j < J; ++j) {
            float* dst = P.data() + ((size_t)c * J + j) * blk;
            for (int s = 0; s < NSLAB; ++s)
                for (int m = 0; m < 128; ++m)
                    for (int i = 0; i < 4; ++i) {
                        const int g = m / Cout, co = m % Cout, k = G * j + g, ci = c * KC + 4 * s + i;   // row = group * Cout + channel
                        float v = (ci < Cin && k < K) ? W[((size_t)co * Cin + ci) * K + k] : 0.f;
                        uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u;
                        float hi; memcpy(&hi, &u, 4);
                        dst[((size_t)s * 128 + m) * 4 + i] = hi;
                        dst[((size_t)(NSLAB + s) * 128 + m) * 4 + i] = v - hi;
                    }
        }
    return P;
}

static int run_case(int B, int C, int T, int K, int dil, int with_res, int accum, int iters) {
    const int Cout = C, pad = (K * dil - dil) / 2;
    const bool v3 = getenv("TC_V3") != nullptr;
    const int G = (v3 && getenv("TC_G") && (C == 32 || C == 64)) ? 128 / C : 1;
    const int N = v3 ? 128 : (Cout > 128 ? 128 : Cout);
    printf("case B=%d C=%d T=%d K=%d dil=%d res=%d accum=%d N=%d: ", B, C, T, K, dil, with_res, accum, N);
    fflush(stdout);
    srand(1234 + C + K);
    std::vector<float> hx((size_t)B * C * T), hw((size_t)Cout * C * K), hb(Cout), hr((size_t)B * Cout * T), hy0((size_t)B * Cout * T);
    for (auto& v : hx) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : hw) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    for (auto& v : hb) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : hr) v = (rand() / (float)RAND_MAX - 0.5f);
    for (auto& v : hy0) v = (rand() / (float)RAND_MAX - 0.5f);
    std::vector<float> hp = G > 1 ? pack_grouped(hw, Cout, C, K, G) : pack_tc(hw, Cout, C, K, N);
    float *dx, *dw, *dp, *db, *dr, *dy, *dyr; int* derr;
    CK(cudaMalloc(&dx, hx.size() * 4)); CK(cudaMalloc(&dw, hw.size() * 4)); CK(cudaMalloc(&dp, hp.size() * 4));
    CK(cudaMalloc(&db, hb.size() * 4)); CK(cudaMalloc(&dr, hr.size() * 4)); CK(cudaMalloc(&dy, hy0.size() * 4));
    CK(cudaMalloc(&dyr, hy0.size() * 4)); CK(cudaMalloc(&derr, 4)); CK(cudaMemset(derr, 0, 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dp, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dr, hr.data(), hr.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dy, hy0.data(), hy0.size() * 4, cudaMemcpyHostToDevice));
    const float slope = 0.1f, scale = accum ? 1.f : 1.f;
    ref_conv<<<dim3((T + 127) / 128, Cout, B), 128>>>(dx, dw, db, with_res ? dr : nullptr, dy, dyr, C, Cout, T, K, dil, pad, slope, scale, accum);
    CK(cudaDeviceSynchronize());
    TcArgs a; memset(&a, 0, sizeof(a));
    a.x = dx; a.x_bs = (long long)C * T; a.x_cs = T; a.Tin = T; a.in_slope = slope;
    a.w = dp; a.bias = db; a.Cin = C; a.K = K; a.dil = dil; a.pad = pad; a.Rows = Cout; a.N = N;
    a.y = dy; a.y_bs = (long long)Cout * T; a.y_cs = T; a.Tout = T;
    if (with_res) { a.res = dr; a.res_bs = (long long)Cout * T; a.res_cs = T; }
    a.scale = scale; a.post_div = 1.f; a.accum = accum; a.err = derr;
    a.sleep_ns = getenv("TC_SLEEP") ? atoi(getenv("TC_SLEEP")) : 0;
    a.dbg = getenv("TC_DBG") ? atoi(getenv("TC_DBG")) : 0;
    a.rows_pad = (TT + (K - 1) * dil + 7) / 8 * 8;
    const size_t smem = smem_bytes(N, a.rows_pad);
    CK(cudaFuncSetAttribute(conv1d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    dim3 grid((T + TT - 1) / TT, (Cout + N - 1) / N, B);
    const bool v2 = getenv("TC_V2") != nullptr;
    b200tts::tc3::Tc3Args a3; memset(&a3, 0, sizeof(a3));
    b200tts::tc2::Tc2Args a2; memset(&a2, 0, sizeof(a2));
    size_t smem2 = 0; dim3 grid2(1);
    if (v3) {
        using namespace b200tts::tc3;
        a3.x = dx; a3.x_bs = a.x_bs; a3.x_cs = T; a3.Tin = T; a3.in_slope = slope; a3.w = dp; a3.bias = db;
        a3.Cin = C; a3.K = K; a3.dil = dil; a3.pad = pad; a3.Rows = Cout; a3.N = 128;
        a3.y = dy; a3.y_bs = a.y_bs; a3.y_cs = T; a3.Tout = T; a3.ups = 1; a3.Tq = T;
        a3.res = a.res; a3.res_bs = a.res_bs; a3.res_cs = a.res_cs; a3.scale = scale; a3.post_div = 1.f; a3.accum = accum;
        a3.rows_pad = a.rows_pad; a3.raw_w = a.rows_pad + 4; a3.B = B; a3.n_ttiles = (T + 255) / 256; a3.n_rtiles = (Cout + 127) / 128;
        a3.err = derr; a3.dbg = a.dbg;
        a3.KJ = K; a3.dil_blk = dil; a3.tstep = 256;
        if (G > 1) {
            const int J = (K + G - 1) / G;
            a3.KJ = J; a3.dil_blk = G * dil; a3.tstep = TSTEP_GROUPED;
            a3.rows_pad = (256 + (J - 1) * G * dil + 7) / 8 * 8; a3.raw_w = a3.rows_pad + 4;
            a3.n_ttiles = (T + a3.tstep - 1) / a3.tstep; a3.n_rtiles = 1;
        }
        smem2 = smem_bytes3(a3.rows_pad, a3.raw_w);
        if (G > 1) { smem2 = (smem2 + 127) / 128 * 128; a3.stage_off = (int)smem2; smem2 += GROUP_XCHG_BYTES; }
        else if (!getenv("TC_STAGE")) { smem2 = (smem2 + 127) / 128 * 128; a3.stage_off = (int)smem2; smem2 += LEAN_STAGE_BYTES; }
        const bool staged = getenv("TC_STAGE") && G == 1;
        if (staged) { a3.stage = 1; a3.stage_off = (int)((smem2 + 15) / 16 * 16); smem2 = (size_t)a3.stage_off + STAGE_BYTES; CK(cudaFuncSetAttribute(conv1d_tc3s_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); }
        if (G > 1) CK(cudaFuncSetAttribute(grouped_kernel(G), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        CK(cudaFuncSetAttribute(conv1d_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int tiles = a3.B * a3.n_ttiles * a3.n_rtiles;
        grid2 = dim3(tiles < sms ? tiles : sms);
        if (G > 1) grouped_kernel(G)<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3);
        else if (staged) conv1d_tc3s_kernel<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3);
        else conv1d_tc3_kernel<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3);
    } else if (v2) {
        using namespace b200tts::tc2;
        a2.x = dx; a2.x_bs = a.x_bs; a2.x_cs = T; a2.Tin = T; a2.in_slope = slope; a2.w = dp; a2.bias = db;
        a2.Cin = C; a2.K = K; a2.dil = dil; a2.pad = pad; a2.Rows = Cout; a2.N = N;
        a2.y = dy; a2.y_bs = a.y_bs; a2.y_cs = T; a2.Tout = T; a2.ups = 1; a2.Tq = T;
        a2.res = a.res; a2.res_bs = a.res_bs; a2.res_cs = a.res_cs; a2.scale = scale; a2.post_div = 1.f; a2.accum = accum;
        a2.rows_pad = a.rows_pad; a2.raw_w = a.rows_pad + 4; a2.B = B; a2.n_ttiles = (T + TT2 - 1) / TT2; a2.n_rtiles = (Cout + N - 1) / N;
        a2.err = derr; a2.tg = taps_per_slot(N);
        smem2 = smem_bytes2(N, a2.rows_pad, a2.raw_w);
        CK(cudaFuncSetAttribute(conv1d_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
        const int tiles = a2.B * a2.n_ttiles * a2.n_rtiles;
        grid2 = dim3(tiles < sms ? tiles : sms);
        conv1d_tc2_kernel<<<grid2, NTHREADS2, smem2>>>(a2);
    } else {
        conv1d_tc_kernel<<<grid, NTHREADS, smem>>>(a);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("KERNEL FAILED: %s\n", cudaGetErrorString(e)); return 1; }
    int herr = 0; CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    std::vector<float> hy(hy0.size()), hyr(hy0.size());
    CK(cudaMemcpy(hy.data(), dy, hy.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hyr.data(), dyr, hy.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0, sumsq = 0;
    for (size_t i = 0; i < hy.size(); ++i) {
        double d = fabs((double)hy[i] - (double)hyr[i]);
        if (d > maxerr) maxerr = d;
        if (fabs(hyr[i]) > maxref) maxref = fabs(hyr[i]);
        sumsq += d * d;
    }
    if (G > 1) printf("grouped G=%d ", G);
    if (a3.stage) printf("staged ");
    printf("%s smem=%zu err_flag=%d max_err=%.3e rms_err=%.3e max_ref=%.3f  %s", v3 ? "v3" : (v2 ? "v2" : "v1"), (v2 || v3) ? smem2 : smem, herr, maxerr, sqrt(sumsq / hy.size()), maxref,
           (herr == 0 && maxerr < 1e-4 * (maxref + 1)) ? "OK" : "MISMATCH");
    if (getenv("TC_TRACE") && v3) {
        unsigned long long* dtr; CK(cudaMalloc(&dtr, (size_t)grid2.x * 32 * 8)); CK(cudaMemset(dtr, 0, (size_t)grid2.x * 32 * 8));
        a3.trace = dtr;
        if (G > 1) b200tts::tc3::grouped_kernel(G)<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3);
        else b200tts::tc3::conv1d_tc3_kernel<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> tr((size_t)grid2.x * 32);
        CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
        printf("\nv3 trace (us since CTA start): prod: chunk0 tile0 tile1 tile3 end | mma: tile0 tile1 tile3 end | epi: t0[start,end] t1[start,end] t3[start,end] end\n");
        for (int blk : {0, (int)grid2.x - 1}) {
            const unsigned long long* r = &tr[(size_t)blk * 32];
            auto us = [&](int k) { return r[k] ? (double)(r[k] - r[0]) / 1e3 : -1.0; };
            printf("  cta %3d: %6.1f %6.1f %6.1f %6.1f %7.1f | %6.1f %6.1f %6.1f %7.1f | [%6.1f %6.1f] [%6.1f %6.1f] [%6.1f %6.1f] %7.1f\n", blk,
                   us(1), us(2), us(3), us(4), us(5), us(8), us(9), us(10), us(11), us(16), us(17), us(18), us(19), us(20), us(21), us(22));
        }
        a3.trace = nullptr;
    } else if (getenv("TC_TRACE") && v2) {
        unsigned long long* dtr; CK(cudaMalloc(&dtr, (size_t)grid2.x * 32 * 8)); CK(cudaMemset(dtr, 0, (size_t)grid2.x * 32 * 8));
        a2.trace = dtr;
        b200tts::tc2::conv1d_tc2_kernel<<<grid2, b200tts::tc2::NTHREADS2, smem2>>>(a2);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> tr((size_t)grid2.x * 32);
        CK(cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost));
        printf("\nv2 trace (us since CTA start): prod: chunk0 tile0 tile1 tile3 end | mma: tile0 tile1 tile3 end | epi: t0[start,end] t1[start,end] t3[start,end] end\n");
        for (int blk : {0, 1, (int)grid2.x / 2, (int)grid2.x - 1}) {
            const unsigned long long* r = &tr[(size_t)blk * 32];
            auto us = [&](int k) { return r[k] ? (double)(r[k] - r[0]) / 1e3 : -1.0; };
            printf("  cta %3d: %6.1f %6.1f %6.1f %6.1f %7.1f | %6.1f %6.1f %6.1f %7.1f | [%6.1f %6.1f] [%6.1f %6.1f] [%6.1f %6.1f] %7.1f\n", blk,
                   us(1), us(2), us(3), us(4), us(5), us(8), us(9), us(10), us(11), us(16), us(17), us(18), us(19), us(20), us(21), us(22));
        }
        a2.trace = nullptr;
    } else if (getenv("TC_TRACE")) {
        const size_t nb = (size_t)grid.x * grid.y * grid.z;
        unsigned long long* dtr; CK(cudaMalloc(&dtr, nb * 16 * 8)); CK(cudaMemset(dtr, 0, nb * 16 * 8));
        a.trace = dtr;
        conv1d_tc_kernel<<<grid, NTHREADS, smem>>>(a);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> tr(nb * 16);
        CK(cudaMemcpy(tr.data(), dtr, nb * 16 * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = ~0ull; for (size_t i = 0; i < nb; ++i) if (tr[i * 16 + 11] && tr[i * 16 + 11] < t0) t0 = tr[i * 16 + 11];
        printf("\ntrace (us since first CTA start): blk start alloc_done prodA0 prodB0 prodAend prodBend acc_full epi_end mma_start mma_firstA mma_end dealloc\n");
        for (size_t i : {(size_t)0, (size_t)1, nb / 3, nb / 2, nb - 2, nb - 1}) {
            const unsigned long long* r = &tr[i * 16];
            auto us = [&](int k) { return r[k] ? (double)(r[k] - t0) / 1e3 : -1.0; };
            printf("  blk %5zu: %8.1f %8.1f | %8.1f %8.1f %8.1f %8.1f | %8.1f %8.1f | %8.1f %8.1f %8.1f | %8.1f || c2: issue %8.1f data %8.1f stored %8.1f arrived %8.1f\n", i, us(11), us(0), us(1), us(2), us(3), us(4), us(5), us(6), us(7), us(8), us(9), us(10), us(12), us(13), us(14), us(15));
        }
        a.trace = nullptr;
    }
    if (iters > 0 && herr == 0 && !accum) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) { if (v3 && G > 1) b200tts::tc3::grouped_kernel(G)<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3); else if (v3 && a3.stage) b200tts::tc3::conv1d_tc3s_kernel<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3); else if (v3) b200tts::tc3::conv1d_tc3_kernel<<<grid2, b200tts::tc3::NTHREADS2, smem2>>>(a3); else if (v2) b200tts::tc2::conv1d_tc2_kernel<<<grid2, b200tts::tc2::NTHREADS2, smem2>>>(a2); else conv1d_tc_kernel<<<grid, NTHREADS, smem>>>(a); }
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
        printf("  %.3f ms  %.1f TFLOP/s (algorithmic fp32)", ms, 2.0 * B * Cout * (double)C * K * T / ms / 1e9);
    }
    printf("\n");
    cudaFree(dx); cudaFree(dw); cudaFree(dp); cudaFree(db); cudaFree(dr); cudaFree(dy); cudaFree(dyr); cudaFree(derr);
    return (herr == 0 && maxerr < 1e-4 * (maxref + 1)) ? 0 : 1;
}

int main(int argc, char** argv) {
    int fails = 0;
    if (argc >= 8 && !strcmp(argv[1], "one"))
        return run_case(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]), atoi(argv[6]), argc >= 9 ? atoi(argv[8]) : 1,
                        argc >= 10 ? atoi(argv[9]) : 0, atoi(argv[7]));   // one B C T K dil iters [res] [accum]
    fails += run_case(1, 32, 256, 1, 1, 0, 0, 0);      // smallest: one chunk group, one tap
    fails += run_case(1, 32, 300, 3, 1, 0, 0, 0);
    fails += run_case(2, 128, 700, 11, 5, 1, 0, 0);
    fails += run_case(2, 256, 300, 3, 3, 1, 1, 0);
    fails += run_case(2, 64, 1000, 7, 3, 0, 0, 0);
    fails += run_case(2, 64, 1004, 11, 5, 1, 1, 0);
    fails += run_case(3, 32, 996, 11, 5, 1, 0, 0);
    fails += run_case(2, 32, 2000, 7, 1, 1, 1, 0);
    fails += run_case(2, 32, 480, 3, 3, 0, 0, 0);
    fails += run_case(2, 32, 724, 3, 5, 1, 0, 0);
    fails += run_case(2, 64, 724, 3, 3, 1, 0, 0);
    fails += run_case(2, 64, 500, 7, 1, 0, 1, 0);
    fails += run_case(2, 32, 500, 5, 2, 0, 1, 0);      // generic (runtime dilation) grouped epilogue
    if (fails == 0 || (argc > 1 && !strcmp(argv[1], "time"))) {
        run_case(32, 128, 9600, 11, 5, 1, 0, 5);       // HiFiGAN stage 1 at cfg2
        run_case(32, 128, 9600, 3, 1, 1, 0, 5);
        run_case(32, 256, 1200, 7, 3, 1, 0, 5);
        run_case(32, 64, 19200, 11, 1, 1, 0, 5);
        run_case(32, 64, 19200, 3, 1, 1, 0, 5);
        run_case(32, 32, 38400, 7, 1, 1, 0, 5);
        run_case(32, 32, 38400, 3, 1, 1, 0, 5);
        run_case(32, 32, 38400, 11, 5, 1, 0, 5);
    }
    printf("FAILS=%d\n", fails);
    return fails;
}
