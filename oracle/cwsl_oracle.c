/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the CWSL_DIGI receive
 * front-end, in plain C, used as the parity oracle for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it; the product path never does.
 *
 * Parity status: the reference ships no tests or golden vectors, so this
 * restatement is PINNED against the reference's own headers compiled here
 * (oracle/_ref/libcwsl_ref.so, built by oracle/Makefile from /root/reference/source/SSBD.hpp
 * + LowPass.hpp): tests/test_oracle.py requires bit-identical tables, float audio and int16
 * output, and tests/golden/ holds vectors generated from that build (tests/golden/make_golden.py).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off (strict IEEE single ops, no FMA contraction).
 *
 * Every function cites the reference lines it follows (paths relative to /root/reference).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define CWSL_PI 3.14159265358979323846 /* source/LowPass.hpp:13 */
#define WAVE_SR 12000u                 /* source/CWSL_DIGI.hpp:51 */
#define SSB_BW 6000u                   /* source/CWSL_DIGI.hpp:52 */

typedef struct {
    uint32_t fs;
    uint32_t filt_order;  /* latency*2*Fs/B      source/SSBD.hpp:62 */
    uint32_t block_size;  /* Fs/B/2              source/SSBD.hpp:71 */
    uint32_t num_ws;      /* FiltOrder/BlockSize source/SSBD.hpp:75 */
    float sign;
    float* filter;        /* [filt_order], normalised */
    float* tone;          /* [2*block_size] interleaved re,im */
    float phase_inc[2];
    float raw_tap_sum;
} oracle_tables_t;

/* source/LowPass.hpp:16-35 */
static float* build_lowpass(size_t order, double bandwidth) {
    float* filter = (float*)malloc(order * sizeof(float));
    filter[0] = (float)0.0;
    filter[order / 2] = (float)1.0;
    const double x0 = -1.0 * order / 2;
    for (size_t n = 1; n < order / 2; ++n) {
        const double xPi = (x0 + n) * CWSL_PI * bandwidth;
        const double y = sin(xPi) / xPi * (0.54 - 0.46 * cos(2.0 * CWSL_PI * n / (double)order));
        filter[n] = (float)y;
        filter[order - n] = (float)y;
    }
    return filter;
}

void oracle_tables_free(oracle_tables_t* t) {
    free(t->filter);
    free(t->tone);
    t->filter = t->tone = NULL;
}

/* source/SSBD.hpp:48-83 (ctor) + :97-123 (Tune). Returns 0, or -1 for invalid arguments
 * (the reference throws std::invalid_argument at SSBD.hpp:54-59 and :100-103). */
int oracle_tables_build(oracle_tables_t* t, uint32_t fs, int32_t demod_freq, int is_usb) {
    const size_t Fs = fs, B = SSB_BW, latency = 1u << 3;
    memset(t, 0, sizeof(*t));
    if (0 == B || (Fs / B / 2) * 2 * B != Fs || Fs < 4 * B) return -1;        /* SSBD.hpp:54 */
    const double F = (double)(float)demod_freq;                              /* Instance.cpp:187 */
    if (fabs(F) > Fs / 2) return -1;                                           /* SSBD.hpp:100 */
    if (fabs(F + B * (is_usb ? 1.0 : -1.0)) > Fs / 2) return -1;               /* SSBD.hpp:102 */

    t->fs = fs;
    t->filt_order = (uint32_t)(latency * 2 * Fs / B);
    t->filter = build_lowpass(t->filt_order, B / (double)Fs);                  /* SSBD.hpp:63 */
    float sum = 0.0;                                                           /* SSBD.hpp:66-68 */
    for (size_t n = 0; n < t->filt_order; sum += t->filter[n++]);
    t->raw_tap_sum = sum;
    for (size_t n = 0; n < t->filt_order; t->filter[n++] /= sum);
    t->block_size = (uint32_t)(Fs / B / 2);
    t->num_ws = t->filt_order / t->block_size;

    const float sign = (float)(is_usb ? 1.0 : -1.0);                           /* SSBD.hpp:110 */
    t->sign = sign;
    const float phase_delta = (float)(-2.0 * CWSL_PI * (F + sign * B / 2.0) / (double)Fs); /* :111 */
    t->tone = (float*)malloc(2 * t->block_size * sizeof(float));
    for (size_t n = 0; n < t->block_size; ++n) {                               /* SSBD.hpp:112-113 */
        const float complex z = cexpf(CMPLXF(0.0f, phase_delta * n));
        t->tone[2 * n] = crealf(z);
        t->tone[2 * n + 1] = cimagf(z);
    }
    const float complex pi = cexpf(CMPLXF(0.0f, phase_delta * t->block_size)); /* SSBD.hpp:114 */
    t->phase_inc[0] = crealf(pi);
    t->phase_inc[1] = cimagf(pi);
    return 0;
}

/* The NCO phase sequence phase_k = phase_inc^k by the reference's float recurrence
 * (source/SSBD.hpp:174, `phase *= phase_inc`, libstdc++ complex multiply = (ac-bd, ad+bc)
 * with every operation rounded separately), starting from (1,0) (SSBD.hpp:120). */
void oracle_phase_table(const float phase_inc[2], size_t n, float* table /*2n*/) {
    float pr = 1.0f, pi = 0.0f;
    const float ir = phase_inc[0], ii = phase_inc[1];
    for (size_t k = 0; k < n; ++k) {
        table[2 * k] = pr;
        table[2 * k + 1] = pi;
        const float a = pr * ir, b = pi * ii, c = pr * ii, d = pi * ir;
        pr = a - b;
        pi = c + d;
    }
}

/* How many IQ blocks of iq_len samples the demod loop accepts before the "af buffer full"
 * guard starts dropping them (source/Instance.cpp:268-271). */
size_t oracle_accepted_blocks(size_t n_iq, size_t iq_len, size_t dec_ratio, size_t af_size) {
    size_t wi = 0, acc = 0;
    for (size_t blk = 0; blk + iq_len <= n_iq; blk += iq_len) {
        if (wi + iq_len > af_size - 1) continue;
        wi += iq_len / dec_ratio;
        ++acc;
    }
    return acc;
}

/* Demodulate one slot in GATHER form:
 *   y[b] = sum_{n=0..NumWS-1, k=b-(NumWS-1)+n >= 0} ( sum_{m<BlockSize} (x[BS*k+m]*tone[m])*h[BS*n+m] ) * phase_k
 * accumulated in ascending n from (0,0) -- the order in which the reference's circular
 * workspace receives its terms (source/SSBD.hpp:164-181) -- followed by the Weaver select of
 * Iterate (source/SSBD.hpp:128-137). Writes n_blocks floats (n_blocks multiple of 4). */
void oracle_demod(const oracle_tables_t* t, const float* phase_table, const float* iq /*interleaved*/,
                  size_t n_blocks, float* af) {
    const size_t BS = t->block_size, NW = t->num_ws;
    const float sign = t->sign;
    for (size_t b = 0; b < n_blocks; ++b) {
        float wr = 0.0f, wi = 0.0f;
        for (size_t n = 0; n < NW; ++n) {
            if (b + n < NW - 1) continue;
            const size_t k = b + n - (NW - 1);
            const float* x = iq + 2 * BS * k;
            float sr = 0.0f, si = 0.0f;
            for (size_t m = 0; m < BS; ++m) {
                const float xr = x[2 * m], xi = x[2 * m + 1];
                const float tr = t->tone[2 * m], ti = t->tone[2 * m + 1];
                const float vr = xr * tr - xi * ti;          /* complex*complex */
                const float vi = xr * ti + xi * tr;
                const float h = t->filter[m + n * BS];
                sr += vr * h;                                /* complex*real, complex += */
                si += vi * h;
            }
            const float pr = phase_table[2 * k], pi = phase_table[2 * k + 1];
            wr += sr * pr - si * pi;                         /* workspace += sum*phase */
            wi += sr * pi + si * pr;
        }
        float o;
        switch (b & 3) {                                     /* SSBD.hpp:132-135 */
            case 0: o = +wr; break;
            case 1: o = -wi * sign; break;
            case 2: o = -wr; break;
            default: o = +wi * sign; break;
        }
        af[b] = o;
    }
}

/* source/Instance.cpp:294-338 (whole buffer, zero tail included) */
void oracle_prepare_audio(float* buf, size_t size, float scale, float* max_out, float* factor_out) {
    float maxVal = -FLT_MAX;
    for (size_t k = 0; k < size; ++k) if (buf[k] > maxVal) maxVal = buf[k];
    float minVal = FLT_MAX;
    for (size_t k = 0; k < size; ++k) if (buf[k] < minVal) minVal = buf[k];
    if (fabsf(minVal) > maxVal) maxVal = fabsf(minVal);
    float factor = 32767.0f / (maxVal + 1.0f);   /* AUDIO_CLIP_VAL, source/CWSL_DIGI.hpp:55 */
    factor *= scale;
    for (size_t k = 0; k < size; ++k) buf[k] *= factor;
    if (max_out) *max_out = maxVal;
    if (factor_out) *factor_out = factor;
}

/* source/Instance.cpp:238-241: add 0.5 then truncate toward zero */
void oracle_quantise(const float* buf, size_t size, int16_t* out) {
    for (size_t k = 0; k < size; ++k) out[k] = (int16_t)(buf[k] + 0.5f);
}

/* Whole chain for one decoder, one slot (steady-state reset, source/Instance.cpp:251).
 * Returns write_index (audio samples produced), or (size_t)-1 on invalid tuning. */
size_t oracle_slot_sb(uint32_t fs, int32_t demod_freq, int is_usb, const float* iq, size_t n_iq, size_t iq_len,
                      float scale, size_t af_size, float* af_raw /*nullable, af_size*/, int16_t* out_i16,
                      float* max_out, float* factor_out) {
    oracle_tables_t t;
    if (oracle_tables_build(&t, fs, demod_freq, is_usb) != 0) return (size_t)-1;
    const size_t dec = fs / WAVE_SR;                                            /* Instance.cpp:192 */
    const size_t acc = oracle_accepted_blocks(n_iq, iq_len, dec, af_size);
    const size_t n_blocks = acc * iq_len / t.block_size;
    float* af = (float*)calloc(af_size, sizeof(float));                         /* Instance.cpp:213 */
    float* ph = (float*)malloc((n_blocks ? n_blocks : 1) * 2 * sizeof(float));
    oracle_phase_table(t.phase_inc, n_blocks, ph);
    oracle_demod(&t, ph, iq, n_blocks, af);
    if (af_raw) memcpy(af_raw, af, af_size * sizeof(float));
    oracle_prepare_audio(af, af_size, scale, max_out, factor_out);
    oracle_quantise(af, af_size, out_i16);
    free(ph);
    free(af);
    oracle_tables_free(&t);
    return n_blocks;
}

size_t oracle_slot(uint32_t fs, int32_t demod_freq, const float* iq, size_t n_iq, size_t iq_len,
                   float scale, size_t af_size, float* af_raw, int16_t* out_i16, float* max_out,
                   float* factor_out) {
    return oracle_slot_sb(fs, demod_freq, 1 /* USB, source/CWSL_DIGI.hpp:53 */, iq, n_iq, iq_len, scale, af_size,
                          af_raw, out_i16, max_out, factor_out);
}

/* Flat accessor for ctypes: filter[filt_order], tone[2*block_size], phase_inc[2], raw sum. */
int oracle_tables_flat(uint32_t fs, int32_t demod_freq, int is_usb, float* filter, float* tone,
                       float* phase_inc, float* raw_tap_sum, uint32_t* dims /*[3]*/) {
    oracle_tables_t t;
    if (oracle_tables_build(&t, fs, demod_freq, is_usb) != 0) return -1;
    memcpy(filter, t.filter, t.filt_order * sizeof(float));
    memcpy(tone, t.tone, 2 * t.block_size * sizeof(float));
    phase_inc[0] = t.phase_inc[0];
    phase_inc[1] = t.phase_inc[1];
    *raw_tap_sum = t.raw_tap_sum;
    dims[0] = t.filt_order; dims[1] = t.block_size; dims[2] = t.num_ws;
    oracle_tables_free(&t);
    return 0;
}
