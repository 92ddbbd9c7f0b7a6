// pipe_peaks.cu -- measures the integer / DPX issue rates of one B200 so the
// DPX-ALU roofline in DESIGN.md / bench.py uses a MEASURED denominator
// (MEASURED_PEAKS.json has no integer peak).  Each kernel runs NCHAIN
// independent dependency chains per thread of one instruction kind, every SM
// full (148*k blocks of 1024 threads); rate = lane-ops / elapsed.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks pipe_peaks.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <string>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

constexpr int NCHAIN = 8;
constexpr int ITERS = 4096;

enum Op { IADD, LOP3, SHFT, IMAD, IMADSHL, DP4A, VMAX, VMAX3, VADDMAX, VADDMAXRELU, SETPSEL, PRMTOP,
          PREDOR, VADDMAX16, VADDMAX16RELU, VMAX316, VADD16, VBMAX16, CELL16, CELL16B, CELL16N, MIX_VADDMAX_IMAD, MIX_VADDMAX_DP4A, MIX_LOP_IMAD, MIX_VMAX3_DP4A_IMAD, CELL_TAG, CELL_PRED, NOPS };
static const char* opname[NOPS] = {"IADD3", "LOP3", "SHF", "IMAD", "IMAD.SHL", "IDP.4A", "VIMNMX", "VIMNMX3", "VIADDMNMX",
    "VIADDMNMX.RELU", "ISETP+SEL", "PRMT", "ISETP+@P LOP3", "VIADDMNMX.S16x2", "VIADDMNMX.S16x2.RELU", "VIMNMX3.S16x2",
    "VIADD.16x2 (__vadd2)", "VIMNMX.S16x2+preds (__vibmax_s16x2)", "cell16(PRMT,VIADD.16x2,2xVIADDMNMX.S16x2,LOP3,2xIMAD; 2 cells)",
    "cell16b(PRMT,VIADDMNMX.S16x2,VIMNMX3.S16x2,LOP3,4xIMAD: both adds as 32-bit IMAD on a biased low half; 2 cells)",
    "cell16n(PRMT,2xVIADDMNMX.S16x2,LOP3,2xIMAD: left candidate carried un-added; 2 cells)", "VIADDMNMX+IMAD", "VIADDMNMX+IDP.4A", "LOP3+IMAD",
    "VIMNMX3+IDP+IMAD", "cell(tagged:LOP3,2xVIADDMNMX,IDP,2xIMAD)", "cell(pred:2xVIMNMX,IADD,IDP,2xISETP,2x@P LOP3)"};
// lane-ops counted per inner iteration per chain
static const int opcount[NOPS] = {1,1,1,1,1,1,1,1,1,1,2,1,2,1,1,1,1,1,7,8,6,2,2,2,3,6,8};

template <int OP>
__global__ void __launch_bounds__(1024) k(int* out, int a0, int b0, int c0)
{
    int v[NCHAIN], w[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) { v[i] = threadIdx.x * 7 + i + a0; w[i] = threadIdx.x + i * 3 + b0; }
    const int b = b0, c = c0;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < NCHAIN; ++i) {
            if (OP == IADD) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(b));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(c));
            if (OP == SHFT) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c));
            if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c));
            if (OP == IMADSHL) asm volatile("mad.lo.s32 %0, %0, 4, %1;" : "+r"(v[i]) : "r"(c));
            if (OP == DP4A) asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(v[i]) : "r"(b), "r"(c));
            if (OP == VMAX) asm volatile("max.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(w[i]));
            if (OP == VMAX3) asm volatile("{.reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(v[i]) : "r"(w[i]), "r"(c));
            if (OP == VADDMAX) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(v[i]) : "r"(b), "r"(w[i]));
            if (OP == VADDMAXRELU) asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 t, t, %2; max.s32 %0, t, 0;}" : "+r"(v[i]) : "r"(b), "r"(w[i]));
            if (OP == SETPSEL) asm volatile("{.reg .pred p; setp.ge.s32 p, %0, %1; selp.s32 %0, %0, %2, p;}" : "+r"(v[i]) : "r"(w[i]), "r"(c));
            if (OP == PRMTOP) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(b), "r"(c));
            if (OP == PREDOR) asm volatile("{.reg .pred p; setp.ge.s32 p, %0, %1; @p or.b32 %0, %0, 0x10;}" : "+r"(v[i]) : "r"(w[i]));
            if (OP == VADDMAX16) v[i] = (int)__viaddmax_s16x2((unsigned)v[i], (unsigned)b, (unsigned)w[i]);
            if (OP == VADDMAX16RELU) v[i] = (int)__viaddmax_s16x2_relu((unsigned)v[i], (unsigned)b, (unsigned)w[i]);
            if (OP == VMAX316) v[i] = (int)__vimax3_s16x2((unsigned)v[i], (unsigned)w[i], (unsigned)c);
            if (OP == VADD16) v[i] = (int)__vadd2((unsigned)v[i], (unsigned)b);
            if (OP == VBMAX16) { bool ph, pl; v[i] = (int)__vibmax_s16x2((unsigned)v[i], (unsigned)w[i], &ph, &pl); w[i] += (int)ph + 2 * (int)pl; }
            if (OP == CELL16) {
                // the packed batch cell (sa_batch16.cuh): two cells (one per 16-bit half) per pass
                unsigned s2, cl, m, h, cn;
                asm volatile("prmt.b32 %0, %1, %2, 0x9180;" : "=r"(s2) : "r"(w[(i + 1) % NCHAIN]), "r"(c));      // (an operand that changes: ptxas hoists a loop-invariant PRMT)
                cl = __vadd2((unsigned)v[i], (unsigned)b);
                m = __viaddmax_s16x2((unsigned)w[i], s2, cl);
                h = __viaddmax_s16x2_relu((unsigned)v[i], (unsigned)c, m);
                cn = h & 0xFFFCFFFCu;
                asm volatile("mad.lo.s32 %0, %1, 16, %0;" : "+r"(w[i]) : "r"(h));
                asm volatile("mad.lo.s32 %0, %1, -16, %0;" : "+r"(w[i]) : "r"(cn));
                v[i] = (int)cn;
            }
            if (OP == CELL16B) {
                // the same cell with both constant adds on the FMA pipe (32-bit IMAD; the low half is biased so that it never
                // changes sign and the carry into the high half is a constant) and the zero clamp as a third max operand
                unsigned s2, cl, ct, m, h, cn;
                asm volatile("prmt.b32 %0, %1, %2, 0x9180;" : "=r"(s2) : "r"(w[(i + 1) % NCHAIN]), "r"(c));      // (an operand that changes: ptxas hoists a loop-invariant PRMT)
                asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(cl) : "r"(v[i]), "r"(a0), "r"(b));
                m = __viaddmax_s16x2((unsigned)w[i], s2, cl);
                asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(ct) : "r"(v[i]), "r"(a0), "r"(c));
                h = __vimax3_s16x2(ct, m, (unsigned)b);
                cn = h & 0xFFFCFFFCu;
                asm volatile("mad.lo.s32 %0, %1, 16, %0;" : "+r"(w[i]) : "r"(h));
                asm volatile("mad.lo.s32 %0, %1, -16, %0;" : "+r"(w[i]) : "r"(cn));
                v[i] = (int)cn;
            }
            if (OP == CELL16N) {
                unsigned s2, m, h, cn;
                asm volatile("prmt.b32 %0, %1, %2, 0x9180;" : "=r"(s2) : "r"(w[(i + 1) % NCHAIN]), "r"(c));      // (an operand that changes: ptxas hoists a loop-invariant PRMT)
                m = __viaddmax_s16x2((unsigned)w[i], s2, (unsigned)v[i]);
                h = __viaddmax_s16x2((unsigned)v[i], (unsigned)c, m);
                cn = h & 0xFFFCFFFCu;
                asm volatile("mad.lo.s32 %0, %1, 16, %0;" : "+r"(w[i]) : "r"(h));
                asm volatile("mad.lo.s32 %0, %1, -16, %0;" : "+r"(w[i]) : "r"(cn));
                v[i] = (int)cn;
            }
            if (OP == MIX_VADDMAX_IMAD) {
                asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(v[i]) : "r"(b), "r"(c));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(b), "r"(c));
            }
            if (OP == MIX_VADDMAX_DP4A) {
                asm volatile("{.reg .s32 t; add.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(v[i]) : "r"(b), "r"(c));
                asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(w[i]) : "r"(b), "r"(c));
            }
            if (OP == MIX_LOP_IMAD) {
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(b), "r"(c));
                asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(w[i]) : "r"(b), "r"(c));
            }
            if (OP == MIX_VMAX3_DP4A_IMAD) {
                asm volatile("dp4a.s32.s32 %0, %1, %2, %0;" : "+r"(w[i]) : "r"(b), "r"(c));
                asm volatile("{.reg .s32 t; max.s32 t, %0, %1; max.s32 %0, t, %2;}" : "+r"(v[i]) : "r"(w[i]), "r"(c));
                asm volatile("mad.lo.s32 %0, %0, 4, %1;" : "+r"(w[i]) : "r"(v[i]));
            }
            if (OP == CELL_TAG) {
                // clean = v & ~3 ; x = dp4a ; m = max(clean + KL, x) ; h = max(w + KT, m) ; acc += h*P ; acc -= clean*P
                int clean, x, m;
                asm volatile("and.b32 %0, %1, -4;" : "=r"(clean) : "r"(v[i]));
                asm volatile("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(x) : "r"(b), "r"(c), "r"(w[i]));
                asm volatile("{.reg .s32 t; add.s32 t, %1, %2; max.s32 %0, t, %3;}" : "=r"(m) : "r"(clean), "r"(b), "r"(x));
                asm volatile("{.reg .s32 t; add.s32 t, %1, %2; max.s32 %0, t, %3;}" : "=r"(v[i]) : "r"(w[i]), "r"(c), "r"(m));
                asm volatile("mad.lo.s32 %0, %1, 16, %0;" : "+r"(w[i]) : "r"(v[i]));
                asm volatile("mad.lo.s32 %0, %1, -16, %0;" : "+r"(w[i]) : "r"(clean));
            }
            if (OP == CELL_PRED) {
                int x, G, y;
                asm volatile("dp4a.s32.s32 %0, %1, %2, %3;" : "=r"(x) : "r"(b), "r"(c), "r"(w[i]));
                asm volatile("{.reg .pred p, q; .reg .s32 G, y;\n\t"
                             "max.s32 G, %0, %1; setp.lt.s32 p, %0, %1;\n\t"
                             "sub.s32 y, G, %3;\n\t"
                             "setp.gt.s32 q, %2, y; max.s32 %0, y, %2;\n\t"
                             "@q or.b32 %1, %1, 0x100; @p or.b32 %1, %1, 0x200;}" : "+r"(v[i]), "+r"(w[i]) : "r"(x), "r"(b));
            }
        }
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; ++i) s += v[i] ^ w[i];
    if (s == 0x7fffffff) out[threadIdx.x] = s;
}

template <int OP>
void run(int* d_out, int sms, double clk_mhz)
{
    const int blocks = sms * 2;
    k<OP><<<blocks, 1024>>>(d_out, 1, 3, 5);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(e0));
        k<OP><<<blocks, 1024>>>(d_out, 1, 3, 5);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    const double laneops = double(blocks) * 1024 * NCHAIN * ITERS * opcount[OP];
    const double tops = laneops / (best * 1e-3) / 1e12;
    printf("{\"op\": \"%s\", \"ms\": %.4f, \"lane_ops_per_s_T\": %.3f, \"per_clk_per_sm_at_%.0fMHz\": %.1f, \"instr_per_iter\": %d}\n",
           opname[OP], best, tops, clk_mhz, tops * 1e12 / (clk_mhz * 1e6) / sms, opcount[OP]);
}

int main()
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    const double clk = clk_khz / 1000.0;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_mhz\": %.0f}\n", p.name, p.multiProcessorCount, clk);
    int* d_out; CK(cudaMalloc(&d_out, 4096));
    const int sms = p.multiProcessorCount;
    run<IADD>(d_out, sms, clk); run<LOP3>(d_out, sms, clk); run<SHFT>(d_out, sms, clk); run<IMAD>(d_out, sms, clk);
    run<IMADSHL>(d_out, sms, clk); run<DP4A>(d_out, sms, clk); run<VMAX>(d_out, sms, clk); run<VMAX3>(d_out, sms, clk);
    run<VADDMAX>(d_out, sms, clk); run<VADDMAXRELU>(d_out, sms, clk); run<SETPSEL>(d_out, sms, clk); run<PRMTOP>(d_out, sms, clk);
    run<PREDOR>(d_out, sms, clk); run<VADDMAX16>(d_out, sms, clk); run<VADDMAX16RELU>(d_out, sms, clk); run<VMAX316>(d_out, sms, clk);
    run<VADD16>(d_out, sms, clk); run<VBMAX16>(d_out, sms, clk); run<CELL16>(d_out, sms, clk); run<CELL16B>(d_out, sms, clk); run<CELL16N>(d_out, sms, clk); run<MIX_VADDMAX_IMAD>(d_out, sms, clk); run<MIX_VADDMAX_DP4A>(d_out, sms, clk);
    run<MIX_LOP_IMAD>(d_out, sms, clk); run<MIX_VMAX3_DP4A_IMAD>(d_out, sms, clk); run<CELL_TAG>(d_out, sms, clk); run<CELL_PRED>(d_out, sms, clk);
    return 0;
}
