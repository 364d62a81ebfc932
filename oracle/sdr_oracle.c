/*
 * sdr_oracle.c -- plain-C, single-threaded restatement of SDRReceiver's channelizer
 * hot path. TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; the product (sdrreceiver_b200/) never does.
 *
 * Parity status: PINNED. The reference ships no tests or golden vectors
 * (SURVEY.md section 4), so this file is pinned instead against the reference itself:
 * tests/test_oracle_vs_ref.py runs oracle/_ref (the unmodified reference sources
 * compiled by oracle/Makefile) and this restatement on the same bytes and demands
 * bit-identical int16 / float output, class by class and for whole ini plans;
 * tests/golden/ holds vectors produced by oracle/_ref for boxes without it.
 * Build flags are pinned with the reference's: -O2 -ffp-contract=off, baseline x86-64.
 *
 * Every function cites the reference lines it follows (paths relative to
 * /root/reference). All sample arithmetic is IEEE binary32 unless noted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846264338327950288
#endif

typedef struct { float re, im; } cf32;

/* std::complex<float> product as g++ -O2 emits it without -ffast-math:
 * (ac - bd, ad + bc), every product and sum rounded to float. */
static cf32 cmulf(cf32 x, cf32 y) {
    cf32 r;
    float ac = x.re * y.re, bd = x.im * y.im, ad = x.re * y.im, bc = x.im * y.re;
    r.re = ac - bd;
    r.im = ad + bc;
    return r;
}

/* ------------------------------------------------------------------ */
/* Oscillator -- oscillator.cpp:4-32 (table), 39-50 (tick)             */
/* ------------------------------------------------------------------ */
typedef struct {
    cf32 *queue;
    int length, ptr;
    cf32 vector;
} osc_t;

static void osc_init(osc_t *o, double sample_rate, double frequency) {
    double angle = 2.0 * M_PI * frequency / sample_rate;
    cf32 rot, v;
    int i;
    rot.re = (float)cos(angle);
    rot.im = (float)sin(angle);
    v.re = 1.0f; v.im = 0.0f;
    o->length = (int)sample_rate;
    o->queue = (cf32 *)malloc(sizeof(cf32) * (size_t)o->length);
    for (i = 0; i < o->length; i++) {
        float norm;
        v = cmulf(v, rot);
        norm = 1.95f - (v.re * v.re + v.im * v.im);
        v.re = v.re * norm;
        v.im = v.im * norm;
        o->queue[i] = v;
    }
    o->vector = v;          /* first stream sample is mixed with queue[length-1] */
    o->ptr = 0;
}

static void osc_tick(osc_t *o) {
    o->ptr++;
    if (o->ptr == o->length) o->ptr = 0;
    o->vector = o->queue[o->ptr];
}

static void osc_free(osc_t *o) { free(o->queue); o->queue = 0; }

/* ------------------------------------------------------------------ */
/* 11-tap half-band decimator, one real arm -- jonti/dsp.cpp:96-148     */
/* (FIRUpdateAndProcessHalfBandQueue), 156-160 (FIRUpdateQueue),        */
/* 163-173 (FIRQueueBackToFront); coefficients halfbanddecimator.h:66-79 */
/* ------------------------------------------------------------------ */
#define HB_N 11
static const float hb11[HB_N] = {
    0.0060431029837374152f, 0.0f, -0.049372515458761493f, 0.0f, 0.29332944952052842f, 0.5f,
    0.29332944952052842f, 0.0f, -0.049372515458761493f, 0.0f, 0.0060431029837374152f
};

typedef struct {
    float *queue;   /* HB_N history slots followed by one block */
    int qptr;
} hbarm_t;

static void hbarm_init(hbarm_t *h, int block) {
    h->queue = (float *)calloc((size_t)block + HB_N, sizeof(float));
    h->qptr = HB_N;
}

static float hbarm_push_out(hbarm_t *h, float sig) {
    const float *q;
    float outsum = 0;
    h->queue[h->qptr] = sig;
    h->qptr++;
    q = h->queue + (h->qptr - HB_N);
    outsum += hb11[0] * (q[0] + q[10]) + hb11[2] * (q[2] + q[8]) + hb11[4] * (q[4] + q[6]) + hb11[5] * q[5];
    return outsum;
}

static void hbarm_push(hbarm_t *h, float sig) {
    h->queue[h->qptr] = sig;
    h->qptr++;
}

/* The copy starts one slot early: the newest sample of the block is dropped
 * from the history (dsp.cpp:169). This is observable output and is kept. */
static void hbarm_back_to_front(hbarm_t *h) {
    if (h->qptr >= HB_N)
        memmove(h->queue, h->queue + ((h->qptr - 1) - HB_N), sizeof(float) * HB_N);
    h->qptr = HB_N;
}

typedef struct { hbarm_t i, q; } hb_t;

/* HalfBandDecimator::decimate -- halfbanddecimator.cpp:43-72 */
static void hb_decimate(hb_t *h, const cf32 *in, int n, cf32 *out) {
    int i, step = 0;
    for (i = 0; i < n; ++i) {
        if (i % 2 == 0) {
            out[step].re = hbarm_push_out(&h->i, in[i].re);
            out[step].im = hbarm_push_out(&h->q, in[i].im);
            step++;
        } else {
            hbarm_push(&h->i, in[i].re);
            hbarm_push(&h->q, in[i].im);
        }
    }
    hbarm_back_to_front(&h->i);
    hbarm_back_to_front(&h->q);
}

/* ------------------------------------------------------------------ */
/* FIR ring (newest sample excluded) -- jonti/dsp.cpp:33-71, 150-154     */
/* ------------------------------------------------------------------ */
typedef struct {
    float *points, *buff;
    int n, buffsize, ptr;
} fir_t;

static void fir_init(fir_t *f, int n, const float *taps) {
    f->n = n;
    f->buffsize = n + 1;
    f->points = (float *)malloc(sizeof(float) * (size_t)n);
    memcpy(f->points, taps, sizeof(float) * (size_t)n);
    f->buff = (float *)calloc((size_t)f->buffsize, sizeof(float));
    f->ptr = 0;
}

static float fir_update_and_process(fir_t *f, float sig) {
    int i, tptr;
    float outsum = 0;
    f->buff[f->ptr] = sig;
    f->ptr++; if (f->ptr >= f->buffsize) f->ptr = 0;
    tptr = f->ptr;
    for (i = 0; i < f->n; i++) {
        outsum += f->points[i] * f->buff[tptr];
        tptr++; if (tptr >= f->buffsize) tptr = 0;
    }
    return outsum;
}

static void fir_update(fir_t *f, float sig) {
    f->buff[f->ptr] = sig;
    f->ptr++; f->ptr %= f->buffsize;
}

static void fir_free(fir_t *f) { free(f->points); free(f->buff); f->points = f->buff = 0; }

/* ------------------------------------------------------------------ */
/* FIRHilbert -- jonti/dsp.cpp:184-231 (ring of len, newest included)    */
/* ------------------------------------------------------------------ */
typedef struct {
    float *points, *buff;
    int n, ptr;
} hilb_t;

static void hilbert_points(int len, int Fs, float *points) {
    float *tmp = (float *)malloc(sizeof(float) * (size_t)len);
    float sumofsquares = 0;
    double gain;
    int n;
    for (n = 0; n < len; n++) {
        if (n == len / 2) tmp[n] = 0;
        else tmp[n] = Fs / (M_PI * (n - len / 2)) * (1 - cos(M_PI * (n - len / 2)));
        sumofsquares += tmp[n] * tmp[n];
    }
    /* dsp.cpp:211 `sqrt(sumofsquares)` on a float resolves to the float overload in C++ */
    gain = sqrtf(sumofsquares);
    for (n = 0; n < len; n++) points[n] = tmp[len - n - 1] / gain;
    free(tmp);
}

static void hilb_init(hilb_t *h, int len, int Fs) {
    h->n = len;
    h->points = (float *)malloc(sizeof(float) * (size_t)len);
    h->buff = (float *)calloc((size_t)len, sizeof(float));
    h->ptr = 0;
    hilbert_points(len, Fs, h->points);
}

static double hilb_update_and_process(hilb_t *h, float sig) {
    int i, tptr;
    float outsum = 0;
    h->buff[h->ptr] = sig;
    h->ptr++; if (h->ptr >= h->n) h->ptr = 0;
    tptr = h->ptr;
    for (i = 0; i < h->n; i++) {
        outsum += h->points[i] * h->buff[tptr];
        tptr++; if (tptr >= h->n) tptr = 0;
    }
    return outsum;
}

static void hilb_free(hilb_t *h) { free(h->points); free(h->buff); h->points = h->buff = 0; }

/* DelayThing<float> -- jonti/dsp.h:79-126 */
typedef struct { float *buf; int ptr, sz; } delay_t;
static void delay_init(delay_t *d, int length) {
    d->sz = length + 1;
    d->buf = (float *)calloc((size_t)d->sz, sizeof(float));
    d->ptr = 0;
}
static float delay_update(delay_t *d, float x) {
    d->buf[d->ptr] = x;
    d->ptr++; d->ptr %= d->sz;
    return d->buf[d->ptr];
}

/* ------------------------------------------------------------------ */
/* firfilter::low_pass, Hamming -- gnuradio/firfilter.cpp:64-108,        */
/* compute_ntaps 110-121, hamming 212-220. Returns ntaps, -1 if the      */
/* reference's sanity_check_1f (122-134) would throw.                    */
/* ------------------------------------------------------------------ */
static int low_pass_ntaps(double fs, double tw) {
    int ntaps = (int)(53 * fs / (22.0 * tw));
    if ((ntaps & 1) == 0) ntaps++;
    return ntaps;
}

static int low_pass_hamming(double gain, double fs, double cutoff, double tw, float *taps, int maxn) {
    int ntaps, M, n;
    float *w, Mf;
    double fwT0, fmax;
    if (fs <= 0.0 || cutoff <= 0.0 || cutoff > fs / 2 || tw <= 0) return -1;
    ntaps = low_pass_ntaps(fs, tw);
    if (ntaps > maxn) return ntaps;
    w = (float *)malloc(sizeof(float) * (size_t)ntaps);
    Mf = (float)(ntaps - 1);
    for (n = 0; n < ntaps; n++) w[n] = 0.54 - 0.46 * cos((2 * M_PI * n) / Mf);
    M = (ntaps - 1) / 2;
    fwT0 = 2 * M_PI * cutoff / fs;
    for (n = -M; n <= M; n++) {
        if (n == 0) taps[n + M] = fwT0 / M_PI * w[n + M];
        else taps[n + M] = sin(n * fwT0) / (n * M_PI) * w[n + M];
    }
    fmax = taps[0 + M];
    for (n = 1; n <= M; n++) fmax += 2 * taps[n + M];
    gain /= fmax;
    for (n = 0; n < ntaps; n++) taps[n] *= gain;
    free(w);
    return ntaps;
}

/* ------------------------------------------------------------------ */
/* vfo -- vfo.cpp:60-176 (init), 235-296 (process), 300-387 (demod)      */
/* ------------------------------------------------------------------ */
#define MAX_STAGES 8
typedef struct {
    int Fs, decim, samples_per_buffer, late, filterbw, samples_out, out_rate, is_leaf;
    float gain;
    osc_t osc;
    hb_t hb[MAX_STAGES];
    cf32 *dec[MAX_STAGES + 1];
    int declen[MAX_STAGES + 1];
    fir_t fir_usb, fir_dec_i, fir_dec_q;
    hilb_t hilbert;
    delay_t delay;
    /* growable outputs */
    int16_t *pcm; float *tap; long n_out, cap_out;
    cf32 *mtap; long n_mtap, cap_mtap; int keep_mtap;
} ovfo_t;

static int ipow2(int e) { return 1 << e; }

static void ovfo_init(ovfo_t *v, int Fs, double mixer, int decim, int samples_per_buffer,
                      int late, int filterbw, float gain, int is_leaf) {
    int a, target_rate, samples_out;
    float taps[4096];
    memset(v, 0, sizeof(*v));
    v->Fs = Fs; v->decim = decim; v->samples_per_buffer = samples_per_buffer;
    v->late = late; v->filterbw = filterbw; v->gain = gain; v->is_leaf = is_leaf;
    osc_init(&v->osc, Fs, mixer);
    target_rate = Fs / ipow2(decim);
    samples_out = samples_per_buffer / ipow2(decim);
    if (is_leaf && late > 0) {                                    /* vfo.cpp:67-100 */
        int n;
        target_rate = target_rate / late;
        samples_out = samples_out / late;
        n = low_pass_hamming(2, target_rate * late, target_rate / 2, (double)target_rate / (late - 1), taps, 4096);
        fir_init(&v->fir_dec_i, n, taps);
        fir_init(&v->fir_dec_q, n, taps);
    }
    v->out_rate = target_rate;
    if (is_leaf && filterbw > 0) {                                /* vfo.cpp:106-124 */
        int n = low_pass_hamming(2, target_rate, filterbw, (double)filterbw / 4, taps, 4096);
        fir_init(&v->fir_usb, n, taps);
    }
    for (a = 0; a < decim; a++) {                                 /* vfo.cpp:127-133 */
        int blk = samples_per_buffer / ipow2(a);
        hbarm_init(&v->hb[a].i, blk);
        hbarm_init(&v->hb[a].q, blk);
    }
    delay_init(&v->delay, (125 - 1) / 2);                         /* vfo.cpp:136-137 */
    hilb_init(&v->hilbert, 125, samples_out);
    v->samples_out = samples_out;
    v->declen[0] = samples_per_buffer;                            /* vfo.cpp:150-156 */
    for (a = 1; a <= decim; a++) v->declen[a] = v->declen[a - 1] / 2;
    for (a = 0; a <= decim; a++) v->dec[a] = (cf32 *)calloc((size_t)v->declen[a], sizeof(cf32));
}

static void ovfo_emit(ovfo_t *v, float usb, int idx) {
    (void)idx;
    if (v->n_out == v->cap_out) {
        v->cap_out = v->cap_out ? v->cap_out * 2 : 65536;
        v->pcm = (int16_t *)realloc(v->pcm, sizeof(int16_t) * (size_t)v->cap_out);
        v->tap = (float *)realloc(v->tap, sizeof(float) * (size_t)v->cap_out);
    }
    /* vfo.cpp:328 / 364: float*float, then double*32768.0, then conversion */
    v->pcm[v->n_out] = (short)(usb * v->gain * 32768.0);
    v->tap[v->n_out] = (float)(usb * v->gain * 32768.0);
    v->n_out++;
}

static void ovfo_usb_demod(ovfo_t *v) {                           /* vfo.cpp:300-332 */
    int i, n = v->declen[v->decim];
    const cf32 *x = v->dec[v->decim];
    for (i = 0; i < n; i++) {
        float usb;
        if (v->filterbw > 0)
            usb = fir_update_and_process(&v->fir_usb,
                    delay_update(&v->delay, x[i].re) - hilb_update_and_process(&v->hilbert, x[i].im));
        else
            usb = delay_update(&v->delay, x[i].re) - hilb_update_and_process(&v->hilbert, x[i].im);
        ovfo_emit(v, usb, i);
    }
}

static void ovfo_usb_decimdemod(ovfo_t *v) {                      /* vfo.cpp:334-387 */
    int i, n = v->declen[v->decim], mark = 0, check = 0, discard = v->late - 1;
    const cf32 *x = v->dec[v->decim];
    for (i = 0; i < n; i++) {
        cf32 curr = x[i];
        if (check == 0) {
            float usb;
            curr.re = fir_update_and_process(&v->fir_dec_i, curr.re);
            curr.im = fir_update_and_process(&v->fir_dec_q, curr.im);
            usb = delay_update(&v->delay, curr.re) - hilb_update_and_process(&v->hilbert, curr.im);
            if (v->filterbw > 0) usb = fir_update_and_process(&v->fir_usb, usb);
            ovfo_emit(v, usb, mark);
            mark++;
            check++;
        } else if (check == discard) {
            fir_update(&v->fir_dec_i, curr.re);
            fir_update(&v->fir_dec_q, curr.im);
            check = 0;
        } else {
            fir_update(&v->fir_dec_i, curr.re);
            fir_update(&v->fir_dec_q, curr.im);
            check++;
        }
    }
}

/* mix + half-band cascade -- vfo.cpp:237-251 */
static void ovfo_front(ovfo_t *v, const cf32 *samples) {
    int i, s;
    for (i = 0; i < v->samples_per_buffer; ++i) {
        v->dec[0][i] = cmulf(v->osc.vector, samples[i]);
        osc_tick(&v->osc);
    }
    for (s = 0; s < v->decim; s++) hb_decimate(&v->hb[s], v->dec[s], v->declen[s], v->dec[s + 1]);
    if (v->keep_mtap) {
        long n = v->declen[v->decim];
        if (v->n_mtap + n > v->cap_mtap) {
            v->cap_mtap = (v->n_mtap + n) * 2;
            v->mtap = (cf32 *)realloc(v->mtap, sizeof(cf32) * (size_t)v->cap_mtap);
        }
        memcpy(v->mtap + v->n_mtap, v->dec[v->decim], sizeof(cf32) * (size_t)n);
        v->n_mtap += n;
    }
}

/* ------------------------------------------------------------------ */
/* One receiver = one stream: sdrj::demodData (sdrj.cpp:266-305) + tree  */
/* ------------------------------------------------------------------ */
#define MAX_MAIN 8
#define MAX_SUB 256
typedef struct {
    int Fs, block, dc;
    int n_main, n_sub;
    ovfo_t *mains[MAX_MAIN];
    ovfo_t *subs[MAX_SUB];
    int sub_main[MAX_SUB];
    cf32 avept;
    cf32 *samples;
} orc_t;

orc_t *orc_create(int Fs, int block, int correct_dc) {
    orc_t *o = (orc_t *)calloc(1, sizeof(orc_t));
    o->Fs = Fs; o->block = block; o->dc = correct_dc;
    o->samples = (cf32 *)calloc((size_t)block, sizeof(cf32));
    return o;
}

int orc_add_main(orc_t *o, double mixer, int decim, int keep_tap) {
    ovfo_t *v;
    if (o->n_main >= MAX_MAIN) return -1;
    v = (ovfo_t *)malloc(sizeof(ovfo_t));
    ovfo_init(v, o->Fs, mixer, decim, o->block, 0, 0, 0.0f, 0);
    v->keep_mtap = keep_tap;
    o->mains[o->n_main] = v;
    return o->n_main++;
}

int orc_add_sub(orc_t *o, int main_idx, int Fs, double mixer, int decim, int samples_per_buffer,
                int late, int filterbw, float gain) {
    ovfo_t *v;
    if (o->n_sub >= MAX_SUB || main_idx < 0 || main_idx >= o->n_main) return -1;
    v = (ovfo_t *)malloc(sizeof(ovfo_t));
    ovfo_init(v, Fs, mixer, decim, samples_per_buffer, late, filterbw, gain, 1);
    o->subs[o->n_sub] = v;
    o->sub_main[o->n_sub] = main_idx;
    return o->n_sub++;
}

/* iq: nblocks * block * 2 bytes of interleaved uint8 I,Q */
void orc_process(orc_t *o, const uint8_t *iq, long nblocks) {
    long b;
    int i, m, s;
    const float a = 1.0f - 0.000001f, c = 0.000001f;
    for (b = 0; b < nblocks; b++) {
        const uint8_t *src = iq + b * (long)o->block * 2;
        for (i = 0; i < o->block; ++i) {
            cf32 curr;
            curr.re = (float)((int)src[2 * i] - 127);        /* jonti/sdr.cpp:43-49 */
            curr.im = (float)((int)src[2 * i + 1] - 127);
            if (o->dc) {                                      /* sdrj.cpp:277-283 */
                o->avept.re = o->avept.re * a + c * curr.re;
                o->avept.im = o->avept.im * a + c * curr.im;
                curr.re -= o->avept.re;
                curr.im -= o->avept.im;
            }
            o->samples[i] = curr;
        }
        for (m = 0; m < o->n_main; m++) {
            ovfo_t *mv = o->mains[m];
            ovfo_front(mv, o->samples);
            for (s = 0; s < o->n_sub; s++) {
                ovfo_t *sv = o->subs[s];
                if (o->sub_main[s] != m) continue;
                ovfo_front(sv, mv->dec[mv->decim]);           /* vfo.cpp:253-266 */
                if (sv->late > 0) ovfo_usb_decimdemod(sv);
                else ovfo_usb_demod(sv);
            }
        }
    }
}

long orc_sub_count(const orc_t *o, int s) { return o->subs[s]->n_out; }
const int16_t *orc_sub_pcm(const orc_t *o, int s) { return o->subs[s]->pcm; }
const float *orc_sub_tap(const orc_t *o, int s) { return o->subs[s]->tap; }
int orc_sub_out_rate(const orc_t *o, int s) { return o->subs[s]->out_rate; }
int orc_sub_samples_out(const orc_t *o, int s) { return o->subs[s]->samples_out; }
long orc_main_count(const orc_t *o, int m) { return o->mains[m]->n_mtap; }
const float *orc_main_tap(const orc_t *o, int m) { return (const float *)o->mains[m]->mtap; }
void orc_clear_outputs(orc_t *o) {
    int i;
    for (i = 0; i < o->n_sub; i++) o->subs[i]->n_out = 0;
    for (i = 0; i < o->n_main; i++) o->mains[i]->n_mtap = 0;
}

static void ovfo_free(ovfo_t *v) {
    int a;
    osc_free(&v->osc);
    for (a = 0; a < v->decim; a++) { free(v->hb[a].i.queue); free(v->hb[a].q.queue); }
    for (a = 0; a <= v->decim; a++) free(v->dec[a]);
    if (v->fir_usb.points) fir_free(&v->fir_usb);
    if (v->fir_dec_i.points) { fir_free(&v->fir_dec_i); fir_free(&v->fir_dec_q); }
    hilb_free(&v->hilbert);
    free(v->delay.buf);
    free(v->pcm); free(v->tap); free(v->mtap);
    free(v);
}

void orc_destroy(orc_t *o) {
    int i;
    for (i = 0; i < o->n_main; i++) ovfo_free(o->mains[i]);
    for (i = 0; i < o->n_sub; i++) ovfo_free(o->subs[i]);
    free(o->samples);
    free(o);
}

/* ------------------------------------------------------------------ */
/* Class-level entry points (same shapes as oracle/ref_prims.cpp)        */
/* ------------------------------------------------------------------ */
void orc_oscillator(double fs, double f, float *out_iq, long n) {
    osc_t o;
    long i;
    osc_init(&o, fs, f);
    for (i = 0; i < n; i++) {
        out_iq[2 * i] = o.vector.re;
        out_iq[2 * i + 1] = o.vector.im;
        osc_tick(&o);
    }
    osc_free(&o);
}

/* the raw table queue[0..L-1], as uploaded to the GPU */
int orc_oscillator_table(double fs, double f, float *out_iq, long maxn) {
    osc_t o;
    osc_init(&o, fs, f);
    if (o.length <= maxn) memcpy(out_iq, o.queue, sizeof(cf32) * (size_t)o.length);
    osc_free(&o);
    return (int)fs;
}

void orc_halfband(const float *in_iq, int block, int nblocks, float *out_iq) {
    hb_t h;
    int b;
    hbarm_init(&h.i, block);
    hbarm_init(&h.q, block);
    for (b = 0; b < nblocks; b++)
        hb_decimate(&h, (const cf32 *)in_iq + (long)b * block, block, (cf32 *)out_iq + (long)b * (block / 2));
    free(h.i.queue); free(h.q.queue);
}

/* HalfBandDecimator with any filter length (halfbanddecimator.cpp:4-41; dsp.cpp:96-173): tables for 11,
 * 23 and 51 taps are loaded and summed by the switch in FIRUpdateAndProcessHalfBandQueue; every other
 * length has all-zero points and no case, so the output is 0. Same queue handling as the 11-tap arm. */
static const float hb23_tab[23] = {
    -0.00014987651418332164, 0.0f, 0.0014748633283609852f, 0.0f, -0.0074416944990005314f, 0.0f, 0.026163522731980929f, 0.0f,
    -0.077593699116544707f, 0.0f, 0.30754683719791986f, 0.5f, 0.30754683719791986f, 0.0f, -0.077593699116544707f, 0.0f,
    0.026163522731980929f, 0.0f, -0.0074416944990005314f, 0.0f, 0.0014748633283609852f, 0.0f, -0.00014987651418332164f};
static const float hb51_tab[51] = {
    0.0010175926971811044, 0.0, -0.0013058886799502411, 0.0, 0.0020730260200910026, 0.0, -0.0034255790572079265, 0.0,
    0.005490505092950141, 0.0, -0.008434405740804745, 0.0, 0.012502602797600649, 0.0, -0.01810260996706492, 0.0,
    0.026000146160530365, 0.0, -0.037851497102093665, 0.0, 0.05801218485928863, 0.0, -0.1025751653146947, 0.0,
    0.31684426465520726, 0.499509647157934, 0.3168442646552072, 0.0, -0.10257516531469468, 0.0, 0.05801218485928862, 0.0,
    -0.03785149710209366, 0.0, 0.02600014616053035, 0.0, -0.018102609967064916, 0.0, 0.012502602797600643, 0.0,
    -0.008434405740804745, 0.0, 0.005490505092950138, 0.0, -0.0034255790572079218, 0.0, 0.0020730260200910026, 0.0,
    -0.0013058886799502405, 0.0, 0.0010175926971811044};

static void hbn_arm(const float *points, int N, const float *in, int stride, int block, int nblocks, float *out) {
    float *queue = (float *)calloc((size_t)block + N, sizeof(float));
    int qptr = N, b, i, k, step;
    for (b = 0; b < nblocks; b++) {
        step = 0;
        for (i = 0; i < block; i++) {
            queue[qptr++] = in[(size_t)stride * ((long)b * block + i)];
            if (i % 2 == 0) {                                       /* halfbanddecimator.cpp:49-60 */
                const float *q = queue + (qptr - N);
                float outsum = 0;
                if (N == 11 || N == 23 || N == 51) {
                    float acc = points[0] * (q[0] + q[N - 1]);
                    for (k = 2; k < (N - 1) / 2; k += 2) acc = acc + points[k] * (q[k] + q[N - 1 - k]);
                    acc = acc + points[(N - 1) / 2] * q[(N - 1) / 2];
                    outsum += acc;
                }
                out[(size_t)stride * ((long)b * (block / 2) + step)] = outsum;
                step++;
            }
        }
        if (qptr >= N) memmove(queue, queue + ((qptr - 1) - N), sizeof(float) * N);   /* dsp.cpp:163-173 */
        qptr = N;
    }
    free(queue);
}

void orc_halfband_n(int taps, const float *in_iq, int block, int nblocks, float *out_iq) {
    static float zeros[256];
    const float *pts = taps == 11 ? hb11 : taps == 23 ? hb23_tab : taps == 51 ? hb51_tab : zeros;
    hbn_arm(pts, taps, in_iq, 2, block, nblocks, out_iq);
    hbn_arm(pts, taps, in_iq + 1, 2, block, nblocks, out_iq + 1);
}

void orc_fir(int ntaps, const float *taps, const float *in, long n, int every, float *out) {
    fir_t f;
    long i, m = 0;
    fir_init(&f, ntaps, taps);
    for (i = 0; i < n; i++) {
        if (every <= 1 || i % every == 0) out[m++] = fir_update_and_process(&f, in[i]);
        else fir_update(&f, in[i]);
    }
    fir_free(&f);
}

void orc_hilbert_points(int len, int Fs, float *points) { hilbert_points(len, Fs, points); }

void orc_usb(int len, int Fs, const float *in_iq, long n, float *out) {
    hilb_t h;
    delay_t d;
    long i;
    hilb_init(&h, len, Fs);
    delay_init(&d, (len - 1) / 2);
    for (i = 0; i < n; i++)
        out[i] = delay_update(&d, in_iq[2 * i]) - hilb_update_and_process(&h, in_iq[2 * i + 1]);
    hilb_free(&h);
    free(d.buf);
}

int orc_low_pass(double gain, double fs, double cutoff, double tw, float *taps, int maxn) {
    return low_pass_hamming(gain, fs, cutoff, tw, taps, maxn);
}

/* DC-removal state trace -- sdrj.cpp:277-283. out[j] = avept entering sample j*every
 * (i.e. after samples 0 .. j*every-1), for j = 0 .. n/every - 1. */
/* vfo::compress -- vfo.cpp:389-424. The IQ forwarder of a VFO that has no sub VFOs and is not a
 * USB demodulator: cstyle 1 packs the upper 4 bits of each arm into one byte, anything else
 * sends int8 I,Q pairs. `scalecomp` is an int (vfo.h:114): re/scalecomp is a float division.
 * The reference converts float -> signed char directly (undefined outside -128..127); its x86
 * build truncates to a 32-bit int and keeps the low byte, which is what is written here.
 * Pinned against the unmodified reference through oracle/_ref on plans/FWD_test.ini. */
static signed char to_s8(float v) { return (signed char)(int)v; }
void orc_compress(const float *in_iq, long n, int scalecomp, int cstyle, unsigned char *out) {
    long i;
    if (cstyle == 1) {
        for (i = 0; i < n; i++) {
            signed char real = to_s8((in_iq[2 * i] / scalecomp) * 128);
            signed char imag = to_s8((in_iq[2 * i + 1] / scalecomp) * 128);
            out[i] = (unsigned char)((real & 0xF0) | (imag & 0xF0) >> 4);
        }
    } else {
        for (i = 0; i < n; i++) {
            out[2 * i] = (unsigned char)to_s8(in_iq[2 * i] * 128);
            out[2 * i + 1] = (unsigned char)to_s8(in_iq[2 * i + 1] * 128);
        }
    }
}

/* The `samples` vector of sdrj::demodData (sdrj.cpp:271-294) for a whole stream from its start:
 * byte -> float (jonti/sdr.cpp:43-49) and, if correct_dc, the running-mean removal. It is what
 * the "Main" spectrum is computed from (sdrj.cpp:296-303). Pinned to the unmodified reference
 * through the fftData signal (oracle/ref_harness.cpp --fft Main). */
void orc_input_samples(const uint8_t *iq, long n, int correct_dc, float *out_iq) {
    const float a = 1.0f - 0.000001f, c = 0.000001f;
    cf32 avept;
    long i;
    avept.re = 0; avept.im = 0;
    for (i = 0; i < n; ++i) {
        cf32 curr;
        curr.re = (float)((int)iq[2 * i] - 127);
        curr.im = (float)((int)iq[2 * i + 1] - 127);
        if (correct_dc) {
            avept.re = avept.re * a + c * curr.re;
            avept.im = avept.im * a + c * curr.im;
            curr.re -= avept.re;
            curr.im -= avept.im;
        }
        out_iq[2 * i] = curr.re;
        out_iq[2 * i + 1] = curr.im;
    }
}

void orc_dc_trace(const uint8_t *iq, long n, int every, float *out_iq) {
    const float a = 1.0f - 0.000001f, c = 0.000001f;
    cf32 avept;
    long i;
    avept.re = 0; avept.im = 0;
    for (i = 0; i < n; ++i) {
        if (i % every == 0) { out_iq[2 * (i / every)] = avept.re; out_iq[2 * (i / every) + 1] = avept.im; }
        avept.re = avept.re * a + c * (float)((int)iq[2 * i] - 127);
        avept.im = avept.im * a + c * (float)((int)iq[2 * i + 1] - 127);
    }
}
