// Fused log-mel front end for 16 kHz audio  (replaces WhisperFeatureExtractor._torch_extract_fbank_features,
// HF:models/whisper/feature_extraction_whisper.py:135-164, mel bank :95-103).
//
//   wave (B, L) fp32, zero padded  ->  reflect-pad 200 | frames of 400 @ hop 160 | periodic Hann | |DFT|^2 (201 bins)
//   -> slaney mel (128) -> log10(max(., 1e-10)) -> per-clip max  [kernel 1]
//   -> max(x, clipmax - 8), (x + 4) / 4 -> fp32 (B,128,T) and / or the bf16 im2col matrix conv1 consumes  [kernel 2]
//
// The real 400-point DFT is evaluated as two half-length real contractions using the even / odd symmetry
// of the windowed frame (w[n] = w[400-n]):   Re_k =  sum_{n=0..200} e[n] cos(2 pi k n / 400),
//                                            Im_k = -sum_{n=1..199} o[n] sin(2 pi k n / 400)
// with e[n] = w[n](x[n] + x[400-n]), o[n] = w[n](x[n] - x[400-n]), e[200] = x[200], e[0] = 0.  All fp32 FMA;
// twiddles are generated in double on the host once.
// Second symmetry: cos(2 pi (200-k) n / 400) = (-1)^n cos(2 pi k n / 400) and sin(2 pi (200-k) n / 400) = -(-1)^n sin(2 pi k n / 400),
// so with the sums split over even and odd n (C_e, C_o, S_e, S_o) one pass yields TWO bins:
//   |X_k|^2 = (C_e + C_o)^2 + (S_e + S_o)^2,   |X_{200-k}|^2 = (C_e - C_o)^2 + (S_e - S_o)^2        (k = 0..100)
// i.e. 40.6 instead of 80.4 kFMA per frame.
#include "common.cuh"
#include "tinyaudio_b200.h"

#include <cmath>
#include <mutex>
#include <vector>

namespace {

constexpr int N_FFT = 400, HOP = 160, N_BINS = 201, N_MELS = 128;
constexpr int HALF = 201;                 // n = 0..200
constexpr int KPITCH = 208;               // bins padded to a multiple of 4 (float4 basis loads)
constexpr int FT = 32;                    // frames per CTA: 73 KB of shared memory -> three CTAs per SM overlap their phases
constexpr int FGROUPS = FT / 8;           // frame groups of 8
constexpr int LM_THREADS = 128;           // 4 frame groups x 32 bin groups (26 active: bins 0..103 of the lower half spectrum)
constexpr int KG_ACTIVE = 26;
constexpr int MAXW = 16;                  // max non-zeros of one mel filter (slaney @128 mels: <= 9)
constexpr int PPITCH = 205;
constexpr int SPAN = FT * HOP + N_FFT - HOP;   // samples touched by FT frames (5360)

__device__ float d_cos[HALF * KPITCH];
__device__ float d_sin[HALF * KPITCH];
__device__ float d_win[HALF];
__device__ int d_mel_start[N_MELS];
__device__ int d_mel_len[N_MELS];
__device__ float d_mel_w[N_MELS * MAXW];

std::once_flag g_tables_once;
int g_tables_rc = 0;

double hz_to_mel(double f) { return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4)) : 3.0 * f / 200.0; }
double mel_to_hz(double m) { return m >= 15.0 ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (m - 15.0)) : 200.0 * m / 3.0; }

int upload_tables() {
    std::vector<float> hc(HALF * KPITCH, 0.f), hs(HALF * KPITCH, 0.f), hw(HALF);
    const double PI = 3.14159265358979323846;
    for (int n = 0; n < HALF; ++n) {
        hw[n] = (float)(0.5 - 0.5 * std::cos(2.0 * PI * n / N_FFT));   // torch.hann_window(400), periodic
        for (int k = 0; k < N_BINS; ++k) {
            const int idx = (int)(((long long)n * k) % N_FFT);
            hc[n * KPITCH + k] = (float)std::cos(2.0 * PI * idx / N_FFT);
            hs[n * KPITCH + k] = (float)std::sin(2.0 * PI * idx / N_FFT);
        }
    }
    // slaney mel bank (HF:audio_utils.py mel_filter_bank(201, 128, 0, 8000, 16000, "slaney", "slaney")), float64 -> float32
    std::vector<double> pts(N_MELS + 2);
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0);
    for (int i = 0; i < N_MELS + 2; ++i) pts[i] = mel_to_hz(m0 + (m1 - m0) * i / (N_MELS + 1));
    std::vector<int> st(N_MELS), ln(N_MELS);
    std::vector<float> ww(N_MELS * MAXW, 0.f);
    for (int m = 0; m < N_MELS; ++m) {
        const double enorm = 2.0 / (pts[m + 2] - pts[m]);
        int first = -1, last = -1;
        std::vector<double> col(N_BINS);
        for (int k = 0; k < N_BINS; ++k) {
            const double f = 8000.0 * k / (N_BINS - 1);
            const double down = (f - pts[m]) / (pts[m + 1] - pts[m]);
            const double up = (pts[m + 2] - f) / (pts[m + 2] - pts[m + 1]);
            const double v = std::fmax(0.0, std::fmin(down, up)) * enorm;
            col[k] = v;
            if ((float)v != 0.f) { if (first < 0) first = k; last = k; }
        }
        if (first < 0) { first = 0; last = -1; }
        if (last - first + 1 > MAXW) { ta_set_error("mel filter %d has %d taps (> %d)", m, last - first + 1, MAXW); return -1; }
        st[m] = first; ln[m] = last - first + 1;
        for (int k = first; k <= last; ++k) ww[m * MAXW + (k - first)] = (float)col[k];
    }
    TA_CHECK_CUDA(cudaMemcpyToSymbol(d_cos, hc.data(), hc.size() * 4));
    TA_CHECK_CUDA(cudaMemcpyToSymbol(d_sin, hs.data(), hs.size() * 4));
    TA_CHECK_CUDA(cudaMemcpyToSymbol(d_win, hw.data(), hw.size() * 4));
    TA_CHECK_CUDA(cudaMemcpyToSymbol(d_mel_start, st.data(), st.size() * 4));
    TA_CHECK_CUDA(cudaMemcpyToSymbol(d_mel_len, ln.data(), ln.size() * 4));
    TA_CHECK_CUDA(cudaMemcpyToSymbol(d_mel_w, ww.data(), ww.size() * 4));
    return 0;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void fill_kernel(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// column of frame f (0..FT-1) inside an FT-float row so that each thread's two float4 loads are conflict free
__device__ __forceinline__ int fcol(int f) {
    const int fg = f >> 3, j = f & 7;
    return (j >> 2) * (FT / 2) + fg * 4 + (j & 3);
}

__global__ void __launch_bounds__(LM_THREADS, 3)
logmel_power_kernel(const float* __restrict__ wave, long long ld_wave, int L, int T, float* __restrict__ raw /*[B,128,T]*/,
                    float* __restrict__ clip_max) {
    extern __shared__ __align__(16) float smem_lm[];
    float* sE = smem_lm;                       // [201][64]
    float* sO = sE + HALF * FT;                // [201][FT]
    float* sX = sO + HALF * FT;                // span samples; the power tile [FT][205] later aliases sE (and the head of sO)
    __shared__ float s_red[16];

    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * FT;
    const float* x = wave + (long long)b * ld_wave;

    // ---- stage the sample span with reflect padding (torch.stft center=True, pad_mode="reflect") ----
    const long long i0 = (long long)t0 * HOP - N_FFT / 2;
    for (int i = tid; i < SPAN; i += LM_THREADS) {
        long long j = i0 + i;
        if (j < 0) j = -j;
        if (j >= L) j = 2LL * (L - 1) - j;
        sX[i] = (j >= 0 && j < L) ? __ldg(x + j) : 0.f;
    }
    __syncthreads();
    // ---- folded, windowed frames ----
    // sX is read at f * 160 + n (bank = n mod 32) and sE / sO are written at n * FT + fcol(f) (bank = fcol(f) mod 32): a lane-per-n,
    // diagonal-in-f walk over 32 x 32 tiles keeps the reads conflict free and the writes 2-way (a lane-per-f walk is 32-way)
    {
        const int lane = tid & 31, wrp = tid >> 5;
        constexpr int N_TILES = (HALF + 31) / 32;
        for (int tile = wrp; tile < N_TILES * (FT / 32); tile += LM_THREADS / 32) {
            const int n = (tile % N_TILES) * 32 + lane;
            const int f0 = (tile / N_TILES) * 32;
            if (n < HALF) {
                const float w = d_win[n];
#pragma unroll 4
                for (int r = 0; r < 32; ++r) {
                    const int f = f0 + ((lane + r) & 31);
                    const float a = sX[f * HOP + n];
                    const float c = (n == 0) ? 0.f : sX[f * HOP + N_FFT - n];
                    float e, o;
                    if (n == 0 || n == N_FFT / 2) { e = w * a; o = 0.f; }
                    else { e = w * (a + c); o = w * (a - c); }
                    sE[n * FT + fcol(f)] = e;
                    sO[n * FT + fcol(f)] = o;
                }
            }
        }
    }
    __syncthreads();

    // ---- DFT: thread = (frame group of 8, bin group of 4 of the LOWER half spectrum); even / odd n accumulate separately ----
    const int fg = tid % FGROUPS, kg = tid / FGROUPS;
    float reE[8][4], reO[8][4], imE[8][4], imO[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { reE[i][j] = 0.f; reO[i][j] = 0.f; imE[i][j] = 0.f; imO[i][j] = 0.f; }
    if (kg < KG_ACTIVE) {
        const float4* cb = reinterpret_cast<const float4*>(d_cos) + kg;
        const float4* sb = reinterpret_cast<const float4*>(d_sin) + kg;
        auto step = [&](int n, float (&re)[8][4], float (&im)[8][4]) {
            const float4 c4 = __ldg(cb + n * (KPITCH / 4));
            const float4 s4 = __ldg(sb + n * (KPITCH / 4));
            const float4 ea = *reinterpret_cast<const float4*>(sE + n * FT + fg * 4);
            const float4 eb = *reinterpret_cast<const float4*>(sE + n * FT + FT / 2 + fg * 4);
            const float4 oa = *reinterpret_cast<const float4*>(sO + n * FT + fg * 4);
            const float4 ob = *reinterpret_cast<const float4*>(sO + n * FT + FT / 2 + fg * 4);
            const float ev[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
            const float ov[8] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
            const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
            const float sv[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    re[i][j] = fmaf(ev[i], cv[j], re[i][j]);
                    im[i][j] = fmaf(ov[i], sv[j], im[i][j]);
                }
        };
#pragma unroll 2
        for (int n = 0; n + 1 < HALF; n += 2) {      // n = 0 .. 199 in (even, odd) pairs
            step(n, reE, imE);
            step(n + 1, reO, imO);
        }
        step(HALF - 1, reE, imE);                    // n = 200 (even)
    }
    __syncthreads();   // every thread is done with sE / sO: reuse them for the power tile
    float* sP = sE;
    if (kg < KG_ACTIVE) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = kg * 4 + j;
                if (k <= (N_FFT / 4)) {              // k = 0 .. 100; its mirror bin is 200 - k
                    const float rp = reE[i][j] + reO[i][j], ip = imE[i][j] + imO[i][j];
                    const float rm = reE[i][j] - reO[i][j], im_ = imE[i][j] - imO[i][j];
                    sP[(fg * 8 + i) * PPITCH + k] = rp * rp + ip * ip;
                    if (k < N_FFT / 4) sP[(fg * 8 + i) * PPITCH + (N_FFT / 2 - k)] = rm * rm + im_ * im_;
                }
            }
    }
    __syncthreads();

    // ---- mel + log10 + running max ----
    float vmax = -INFINITY;
    for (int i = tid; i < N_MELS * FT; i += LM_THREADS) {
        const int f = i % FT, m = i / FT;
        const int t = t0 + f;
        if (t < T) {
            const int k0 = d_mel_start[m], len = d_mel_len[m];
            float acc = 0.f;
            for (int j = 0; j < len; ++j) acc = fmaf(d_mel_w[m * MAXW + j], sP[f * PPITCH + k0 + j], acc);
            const float v = log10f(fmaxf(acc, 1e-10f));
            raw[((long long)b * N_MELS + m) * T + t] = v;
            vmax = fmaxf(vmax, v);
        }
    }
    vmax = warp_max(vmax);
    if ((tid & 31) == 0) s_red[tid >> 5] = vmax;
    __syncthreads();
    if (tid == 0) {
        float m = s_red[0];
        for (int w = 1; w < LM_THREADS / 32; ++w) m = fmaxf(m, s_red[w]);
        if (m > -INFINITY) atomic_max_float(clip_max + b, m);
    }
}

// floor + affine; writes fp32 (B,128,T) [optional] and the bf16 conv1 im2col matrix [B*T, 3*128] (tap-major) [optional]
__global__ void logmel_finalize_kernel(const float* __restrict__ raw, const float* __restrict__ clip_max /*or NULL*/,
                                       int T, float* __restrict__ out_f32, bf16* __restrict__ out_im2col) {
    __shared__ float tile[32][N_MELS + 1];   // [frame][mel]
    const int b = blockIdx.y, t0 = blockIdx.x * 32;
    const float floor_v = clip_max ? (clip_max[b] - 8.0f) : -INFINITY;
    for (int i = threadIdx.x; i < N_MELS * 32; i += blockDim.x) {
        const int f = i % 32, m = i / 32;
        const int t = t0 + f;
        float v = 0.f;
        if (t < T) {
            const long long idx = ((long long)b * N_MELS + m) * T + t;
            v = raw[idx];
            if (clip_max) v = (fmaxf(v, floor_v) + 4.0f) * 0.25f;
            if (out_f32) out_f32[idx] = v;
        }
        tile[f][m] = v;
    }
    __syncthreads();
    if (!out_im2col) return;
    // row (b, t) of the im2col matrix = [x(t-1) | x(t) | x(t+1)], zero outside [0, T)
    for (int i = threadIdx.x; i < 32 * N_MELS; i += blockDim.x) {
        const int m = i % N_MELS, f = i / N_MELS;
        const int t = t0 + f;
        if (t >= T) continue;
        const bf16 v = __float2bfloat16_rn(tile[f][m]);
        // x(t) is tap 1 of row t, tap 0 of row t+1, tap 2 of row t-1
        bf16* base = out_im2col + ((long long)b * T) * (3 * N_MELS);
        base[(long long)t * (3 * N_MELS) + N_MELS + m] = v;
        if (t + 1 < T) base[(long long)(t + 1) * (3 * N_MELS) + m] = v;
        if (t - 1 >= 0) base[(long long)(t - 1) * (3 * N_MELS) + 2 * N_MELS + m] = v;
        if (t == 0) base[m] = __float2bfloat16_rn(0.f);                                             // tap 0 of row 0
        if (t == T - 1) base[(long long)t * (3 * N_MELS) + 2 * N_MELS + m] = __float2bfloat16_rn(0.f);   // tap 2 of last row
    }
}

}  // namespace

TA_API int ta_logmel_workspace_floats(int B, int L, long long* n_floats) {
    TA_REQUIRE(n_floats, "null");
    const long long T = L / HOP;
    *n_floats = (long long)B * N_MELS * T + B;
    return 0;
}

TA_API int ta_logmel_fwd(const float* wave, long long ld_wave, int B, int L, float* workspace, float* out_f32,
                         void* out_conv1_im2col_bf16, void* stream) {
    TA_REQUIRE(wave && workspace, "ta_logmel_fwd: null pointer");
    TA_REQUIRE(L >= N_FFT / 2 + 1, "ta_logmel_fwd: clip of %d samples is too short for reflect padding", L);
    std::call_once(g_tables_once, [] { g_tables_rc = upload_tables(); });
    if (g_tables_rc) return g_tables_rc;
    if (B == 0) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int T = L / HOP;   // torch.stft gives 1 + L/HOP frames; the last one is dropped (HF:whisper/fe:151)
    float* raw = workspace;
    float* clip_max = workspace + (long long)B * N_MELS * T;
    fill_kernel<<<(B + 127) / 128, 128, 0, st>>>(clip_max, B, -INFINITY);
    TA_LAUNCH_CHECK();
    static_assert(FT * PPITCH <= 2 * HALF * FT, "power tile must fit inside sE | sO");
    const int smem = (2 * HALF * FT + SPAN) * 4;
    static bool done = false;
    if (!done) {
        TA_CHECK_CUDA(cudaFuncSetAttribute(logmel_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        done = true;
    }
    dim3 grid((T + FT - 1) / FT, B);
    logmel_power_kernel<<<grid, LM_THREADS, smem, st>>>(wave, ld_wave, L, T, raw, clip_max);
    TA_LAUNCH_CHECK();
    if (out_f32 || out_conv1_im2col_bf16) {
        dim3 g2((T + 31) / 32, B);
        logmel_finalize_kernel<<<g2, 256, 0, st>>>(raw, clip_max, T, out_f32, reinterpret_cast<bf16*>(out_conv1_im2col_bf16));
        TA_LAUNCH_CHECK();
    }
    return 0;
}

// (B,128,T) fp32 features computed elsewhere (the reference's CPU collator) -> conv1 im2col matrix
TA_API int ta_mel_to_conv1_im2col(const float* mel, int B, int T, void* out_conv1_im2col_bf16, void* stream) {
    TA_REQUIRE(mel && out_conv1_im2col_bf16, "ta_mel_to_conv1_im2col: null pointer");
    if (B == 0 || T == 0) return 0;
    dim3 g2((T + 31) / 32, B);
    logmel_finalize_kernel<<<g2, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(mel, nullptr, T, nullptr,
                                                                                  reinterpret_cast<bf16*>(out_conv1_im2col_bf16));
    TA_LAUNCH_CHECK();
    return 0;
}
