/* ldd_b200.h -- C ABI of libldd_b200.so: the B200 (sm_100a) implementation of ld-decode's RF
 * demodulation + sync + TBC hot path.
 *
 * The reference (wondras/ld-decode) has no FFI layer for this path: the boundary is the Python
 * class surface of lddecode_core.py (RFDecode.demodblock / demod / audio_phase2, Field.get_syncpeaks,
 * Field.downscale, lddutils.scale / load_packed_data_*).  Each entry point below names the reference
 * code it replaces; INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 *
 * Conventions
 *   - plain C types only; every pointer marked "dev" is a device pointer owned by the caller
 *     (e.g. torch.Tensor.data_ptr()); tables passed to ldd_set_filter are host pointers.
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*) unless noted.
 *   - return 0 on success, negative LDD_E* otherwise; ldd_last_error() gives the text.
 *   - one handle per (GPU, stream user); a handle is not thread-safe.
 *   - there is no CPU fallback: without a CUDA device ldd_create fails with LDD_ECUDA.
 *
 * Plane convention (device, float32, structure of arrays): the reference returns a float64
 * record array in Hz (lddecode_core.py:314-316).  Here each field is its own float32 plane and
 * the two planes that ride on the carrier offset are stored RELATIVE to ire0 to keep sub-Hz
 * resolution in float32:
 *     plane[LDD_P_DEMOD]    = demod    - ire0      (Hz)
 *     plane[LDD_P_DEMOD05]  = demod_05 - ire0      (Hz)
 *     plane[LDD_P_SYNC]     = demod_sync           (0..1)  -- FLOAT64: the greedy peak search compares
 *                                                     neighbouring samples, so it must see the reference's ordering
 *     plane[LDD_P_BURST]    = demod_burst          (Hz, zero-centred)
 *     plane[LDD_P_PILOT]    = demod_pilot          (Hz, zero-centred; PAL only)
 * Analog audio is float64 in absolute Hz, as the reference's rv_audio (lddecode_core.py:322-328).
 */
#ifndef LDD_B200_H
#define LDD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDD_ABI_VERSION 2

#define LDD_SYSTEM_NTSC 0
#define LDD_SYSTEM_PAL 1

/* capture sample formats (lddutils.py:131-229, ddunpack.c) */
#define LDD_FMT_U8 0       /* load_unpacked_data_u8: one byte per sample */
#define LDD_FMT_S16 1      /* load_unpacked_data_s16 */
#define LDD_FMT_U16 2      /* already unpacked 10-bit samples, 0..1023 */
#define LDD_FMT_R30 3      /* 3 x 10 bit in a little-endian u32 (ddpack.c); decoded as load_packed_data_3_32: raw 0..1023 */
#define LDD_FMT_LDS40 4    /* 4 x 10 bit in 5 bytes, MSB first (load_packed_data_4_40) */

/* filter tables built on the host exactly as RFDecode.computefilters does (lddecode_core.py:147-279) */
#define LDD_F_RFVIDEO 0    /* Filters['RFVideo'] * Filters['MTF']**mtf_level, blocklen complex128 */
#define LDD_F_VIDEO 1      /* Filters['FVideo'] */
#define LDD_F_VIDEO05 2    /* Filters['FVideo05'] (the np.roll by F05_offset is applied by the library) */
#define LDD_F_BURST 3      /* Filters['FVideoBurst'] */
#define LDD_F_PILOT 4      /* Filters['FVideoPilot'] (PAL) */
#define LDD_F_AUDIO_L 5    /* Filters['audio_lfilt'], already sliced: 2*blocklen/audio_fdiv1 entries */
#define LDD_F_AUDIO_R 6    /* Filters['audio_rfilt'] */
#define LDD_F_AUDIO_LPF2 7 /* Filters['audio_lpf2'], blocklen/4 entries */
#define LDD_F_MTF 8        /* Filters['MTF'], blocklen complex128.  Once it is set, LDD_F_RFVIDEO is taken as Filters['RFVideo']
                              alone (level 0) and ldd_set_mtf_level() applies MTF**level on the device */

#define LDD_P_DEMOD 0
#define LDD_P_DEMOD05 1
#define LDD_P_SYNC 2
#define LDD_P_BURST 3
#define LDD_P_PILOT 4

#define LDD_PREC_F64 0     /* every transform in float64: the reference-exact lane */
#define LDD_PREC_F32 1     /* float32 shared-memory lane */
#define LDD_PREC_MIXED 2   /* float32 lane, then float64 re-run of every block that holds a demod_05 sample within
                              a guard band (16 Hz, 8x the largest float32 error measured) of a sync threshold: the sync
                              decisions, hence demod_sync and the peak indices, are the float64 lane's */

#define LDD_OK 0
#define LDD_EINVAL (-1)
#define LDD_ESHORT (-2)    /* capture too short for the request; the reference returns None (lddecode_core.py:386-392) */
#define LDD_ECUDA (-3)
#define LDD_ENOMEM (-4)
#define LDD_ECAP (-5)      /* caller buffer too small */

typedef struct ldd_handle ldd_handle;

/* Mirrors the state RFDecode.__init__ derives (lddecode_core.py:120-145, 223-279). */
typedef struct ldd_config {
    int abi_version;        /* LDD_ABI_VERSION */
    int device;             /* CUDA device ordinal */
    int system;             /* LDD_SYSTEM_* */
    int blocklen;           /* RFDecode.blocklen, power of two, 4096..262144 */
    int blockcut;           /* RFDecode.blockcut (1024) */
    int blockcut_end;       /* RFDecode.blockcut_end = Filters['F05_offset'] (32) */
    int f05_offset;         /* Filters['F05_offset'] */
    int precision;          /* LDD_PREC_* */
    int decode_analog_audio;
    int audio_slice_lo;     /* Filters['audio_fdslice_lo'].start */
    int audio_slice_hi;     /* Filters['audio_fdslice_lo'].stop  */
    int linelen;            /* RFDecode.linelen */
    int outlinelen;         /* SysParams['outlinelen'] */
    double freq_hz;         /* RFDecode.freq_hz */
    double freq_arf;        /* Filters['freq_arf'] */
    double audio_lowfreq;   /* Filters['audio_lowfreq'] */
    double ire0;            /* SysParams['ire0'] */
    double hz_ire;          /* SysParams['hz_ire'] */
    double vsync_ire;       /* SysParams['vsync_ire'] */
    double sync_lo_hz;      /* iretohz(-55) (lddecode_core.py:308) */
    double sync_hi_hz;      /* iretohz(-25) */
    double fpsync_b0;       /* butter(1, 0.05/freq_half) numerator/denominator (lddecode_core.py:213) */
    double fpsync_b1;
    double fpsync_a1;
} ldd_config;

int ldd_abi_version(void);
/* number of CUDA devices visible (0 when there is none); never fails */
int ldd_device_count(void);

int ldd_create(const ldd_config* cfg, ldd_handle** out);
void ldd_destroy(ldd_handle* h);
const char* ldd_last_error(ldd_handle* h);

/* Upload one host-built frequency-domain table (interleaved re,im float64, n complex entries).
 * Synchronous (set-up time).  The MTF level is NOT baked into a table: see ldd_set_mtf_level. */
int ldd_set_filter(ldd_handle* h, int id, const double* table, int n);

/* RFVideo * MTF ** mtf_level (lddecode_core.py:290-293) for every later demodulation on `stream`: one small
 * stream-ordered kernel, so following the reference's per-frame level changes on CAV discs (:1300-1306) costs neither
 * a host-side complex power nor a blocking upload.  Needs LDD_F_RFVIDEO (level 0) and LDD_F_MTF. */
int ldd_set_mtf_level(ldd_handle* h, double mtf_level, void* stream);

/* Per-block MTF level for whole-range decodes of CAV discs, where the reference lowers mtf_level by 1e-4 with every
 * frame (Framer.readframe, lddecode_core.py:1300-1306: the level used for a frame is 1 - framenr/10000 of the frame
 * before it).  After this call a block whose kept samples are centred in the n-th frame period after capture sample
 * pos0_sample (n may be negative) is demodulated with level max(L0 + n * step_per_period, 0), L0 being the level of
 * ldd_set_mtf_level -- the level of the frame that starts at pos0_sample.  period_samples = 0 switches the ramp off
 * (every block uses L0, the behaviour of RFDecode.demod).  Blocks centred before capture sample hold_until_sample use
 * hold_level instead: the reference decodes the FIRST frame of a run with its start-up level (Framer.mtf_level = 1,
 * lddecode_core.py:1334) whatever the disc position; pass -1e300 for none.  The correction MTF^(n step) is applied inside
 * the kernel as a third-order series in n*step*log(MTF): keep |n * step| below ~1e-2 (re-base L0 and pos0 per range). */
int ldd_set_mtf_ramp(ldd_handle* h, double pos0_sample, double period_samples, double step_per_period,
                     double hold_until_sample, double hold_level);

/* ---- kernel (1): unpack.  Replaces ddunpack.c:11-36 and lddutils.py:150-229. ------------------ */
/* words[nwords] (LE u32, dev) -> out[3*nwords] int16 = ((field)-512)<<6, exactly ddunpack.c */
int ldd_unpack_r30_ddunpack(const uint32_t* words_dev, size_t nwords, int16_t* out_dev, void* stream);
/* samples [first, first+n) of a packed capture -> int16/uint16 raw 0..1023 (the Python loaders) or float32 */
int ldd_unpack_raw(const void* src_dev, int fmt, size_t first_sample, size_t n, uint16_t* out_dev, void* stream);
int ldd_unpack_f32(const void* src_dev, int fmt, size_t first_sample, size_t n, float* out_dev, void* stream);

/* ---- kernels (2)+(3)+(4a): block demodulation.  Replaces RFDecode.demodblock + the stitching of
 * RFDecode.demod (lddecode_core.py:288-330, 373-427) and lddutils.unwrap_hilbert (:320-334). ---- */

/* Geometry of demod(start, length) as the reference computes it (lddecode_core.py:374-385, 400, 417). */
typedef struct ldd_range {
    long long first_sample;   /* capture index of output sample 0 (start - blockcut, or 0) */
    long long nblocks;
    long long total_out;      /* len(output) */
    long long audio1_len;     /* len(output_audio) before phase 2; 0 when audio is off */
    long long audio2_len;     /* len(audio_phase2(output_audio)) */
    long long last_needed;    /* one past the last capture sample any block reads */
} ldd_range;
int ldd_demod_range_query(ldd_handle* h, long long start, long long length, ldd_range* out);

/* Demodulate demod(start, length).  rf_dev holds the capture (or a window of it) in format fmt;
 * rf_base is the capture sample index of rf_dev[0] and rf_len the number of samples available
 * from there.  planes_dev[LDD_P_*] each hold r.total_out float32 (float64 for LDD_P_SYNC); audio1_* hold r.audio1_len
 * doubles (may be NULL when audio is off).  Returns LDD_ESHORT when a block would read past
 * rf_base+rf_len (the reference's loader returns None there). */
int ldd_demod_range(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                    long long start, long long length,
                    void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev, void* stream);

/* The same kernel on an explicit block grid: nblocks blocks, block j reads capture samples
 * [first_sample + j*stride, +blocklen) and contributes output samples [j*stride, j*stride+copylen). */
int ldd_demod_blocks(ldd_handle* h, const void* rf_dev, int fmt, long long rf_base, long long rf_len,
                     long long first_sample, long long nblocks, long long total_out,
                     void* const* planes_dev, double* audio1_l_dev, double* audio1_r_dev,
                     long long audio1_len, void* stream);

/* RFDecode.demodblock (lddecode_core.py:288-330) on one block: rf_dev holds >= blocklen samples;
 * every plane receives all blocklen samples (nothing cut), audio_* receive 2*blocklen/audio_fdiv1
 * float64 samples (may be NULL). */
int ldd_demodblock(ldd_handle* h, const void* rf_dev, int fmt, long long rf_len,
                   void* const* planes_dev, double* audio_l_dev, double* audio_r_dev, void* stream);

/* LDD_PREC_MIXED only, synchronous: how many blocks of the last ldd_demod_* call were re-run in float64. */
int ldd_mixed_stats(ldd_handle* h, long long* flagged_blocks, long long* total_blocks);

/* RFDecode.audio_phase2 (lddecode_core.py:335-371): in[len] -> out[len/4], float64, dev. */
int ldd_audio_phase2(ldd_handle* h, const double* in_l_dev, const double* in_r_dev, long long len,
                     double* out_l_dev, double* out_r_dev, void* stream);

/* downscale_audio (lddecode_core.py:431-484), batched over fields: 48 kHz (or any rate) int16 L/R PCM resampled along the
 * line positions from the phase-2 audio.  Output sample i of field f belongs to time arange[i] of the reference's
 * np.arange(timeoffset, frametime + 1/freq, 1/freq): t0 = arange[0], t1 = arange[1] (numpy fills t0 + i * (t1 - t0) from
 * i = 2 on), nout = len(arange) - 1, computed by the caller, who also carries arange[-1] - frametime into the next field.  linelocs [nfields][ll_stride]
 * (+ lineloc_add) is Field.linelocs, nll its length (linecount + 4), audio_base_dev[f] the index of the field window's
 * first audio sample in audio_*_dev (NULL: 0), scale the reference's `scale` argument (64).  out_dev receives
 * 2 * nout[f] int16 at out_off[f].  status bit 16: an index left the arrays (the reference raises). */
int ldd_downscale_audio(ldd_handle* h, const double* audio_l_dev, const double* audio_r_dev, long long audio_len,
                        const long long* audio_base_dev, const double* linelocs_dev, int ll_stride,
                        const int* nll_dev, const double* t0_dev, const double* t1_dev, const int* nout_dev,
                        const long long* out_off_dev, int nfields, int max_nout, double lineloc_add, double scale,
                        double line_period_us, double audio_lfreq, double audio_rfreq,
                        short* out_dev, int* status_dev, void* stream);

/* ---- kernel (4): sync-pulse peak list.  Bit-exact Field.get_syncpeaks (lddecode_core.py:497-516)
 * over sync_dev[0..n) (the float64 demod_sync plane) starting at index `start`, with the handle's
 * linelen.  peaks_dev/vals_dev receive up to `cap` peak indices and ds[peak] values; count_dev[0]
 * receives the number of peaks found (if > cap the list is truncated: call again with a larger
 * buffer), count_dev[1] the number of chase steps the stitching pass had to take itself. */
int ldd_sync_peaks(ldd_handle* h, const double* sync_dev, long long n, long long start,
                   long long* peaks_dev, double* vals_dev, int cap, int* count_dev, void* stream);

/* HOST function, same algorithm and result as ldd_sync_peaks for samples that are already in host
 * memory (sync_host[0..n)); used by the field walk for the short prefix of a window that does not
 * start on a peak of the capture-wide chase.  count receives the number of peaks found (> cap: truncated). */
int ldd_sync_peaks_host(ldd_handle* h, const double* sync_host, long long n, long long start,
                        long long* peaks, double* vals, int cap, int* count);

/* Small transfers done by a kernel instead of the copy engines (page-locked host memory is device-addressable):
 * the per-field tables going up and the peak list coming down are a few hundred KB each, but on the copy engines
 * they queue behind the bulk capture upload / field download of a streaming decode (0.7 ms each at 36 MB).
 * ldd_copy_small: nbytes from src to dst, either of which may be page-locked host memory or device memory.
 * ldd_peaks_to_host: count_dev[0..1] and the first min(count_dev[0], cap) entries of the peak list. */
int ldd_copy_small(void* dst, const void* src, size_t nbytes, void* stream);
int ldd_peaks_to_host(const long long* peaks_dev, const double* vals_dev, const int* count_dev, int cap,
                      long long* peaks_host, double* vals_host, int* count_host, void* stream);

/* HOST function: the peak list of Field.get_syncpeaks for the window [b, b + wl) of a plane, cut out of the peak
 * list gpeaks[0..ngpeaks) of a chase that covered the plane from sample gstart while i < gend (lddecode_core.py:497-516
 * run once per capture instead of once per 1e6-sample read).  The window may start on a global peak or anywhere else:
 * its own chase is reconstructed from what the global chase is known to have seen (empty half-line windows between
 * peaks, arg-max windows at peaks).  Returns 1 and the index range [*k0, *k1) of gpeaks that is the window's list
 * (positions relative to the plane; subtract b), or 0 when the samples themselves are needed to decide. */
int ldd_window_peaks_from_global(const long long* gpeaks, int ngpeaks, long long gstart, long long gend,
                                 long long b, long long wl, int linelen, int* k0, int* k1);

/* ---- field location (lddecode_core.py:518-787, 889-957, 962-1021, 1054-1133) -------------------- */
#define LDD_FIELD_NOVSYNC 0    /* len(vsyncs) == 0: not a field, nextfieldoffset = start + 200 lines          */
#define LDD_FIELD_SHORT 1      /* one vsync / too few peaks after the second: jump, not valid                 */
#define LDD_FIELD_LOCATED 2    /* linelocs1 / linebad produced                                                */
#define LDD_FIELD_BADLINES 3   /* compute_linelocs raised in the reference: field invalid                     */
#define LDD_FIELD_CRASH 4      /* the reference itself raises out of Field.__init__ (vsync in the first 11 peaks) */

typedef struct ldd_field {
    int stage;                 /* LDD_FIELD_* */
    int istop;
    int linecount;             /* 262/263 NTSC, 312/313 PAL */
    int npeaks;
    int nvsyncs;
    int vsyncs[4][3];          /* first four rows of Field.vsyncs: (peak index, line0 peak index, istop) */
    long long nextfieldoffset; /* Field.nextfieldoffset */
    long long tbcstart;        /* Field.tbcstart */
    double med_hsync;
    double hsync_tolerance;
} ldd_field;

/* HOST function (no device work): Field.determine_vsyncs / determine_field / compute_linelocs and the
 * early-outs of Field.__init__ on one field window.  peaks/vals are HOST arrays: the window's peak
 * list (indices relative to the window) and demod_sync at those peaks; window_len = len(demod_sync)
 * of the window; start = Field.start.  linelocs1/linebad (HOST, ll_cap >= linecount+4 entries)
 * receive Field.linelocs1 / Field.linebad when stage == LDD_FIELD_LOCATED. */
int ldd_field_locate(ldd_handle* h, const long long* peaks, const double* vals, int npeaks,
                     long long window_len, long long start, ldd_field* out,
                     double* linelocs1, unsigned char* linebad, int ll_cap);

/* HOST function: Field.determine_field(peaknum) (lddecode_core.py:544-588) on its own, for callers that step through the
 * reference's methods one by one: *line0 = peak index of the last regular hsync before peaknum (-1: None, also for
 * peaknum < 11 where the reference returns None), *vote = the parity vote.  med_hsync / hsync_tolerance as returned by
 * ldd_field_locate (Field.get_hsync_median). */
int ldd_field_vote(ldd_handle* h, const long long* peaks, const double* vals, int npeaks, long long window_len,
                   double med_hsync, double hsync_tolerance, int peaknum, int* line0, int* vote);

/* HOST function: the field-to-field walk of Framer.readfield (lddecode_core.py:1194-1223) over planes
 * that were demodulated on one block grid (ldd_demod_blocks with first_sample = plane_origin, so that
 * plane index k <-> capture sample plane_origin + blockcut + k).  gpeaks/gvals: the peak list of
 * ldd_sync_peaks over the whole sync plane from 0.  The walk starts at first_readsample and stops when
 * readsample >= stop_readsample, when a window leaves the planes, or when the capture (ncap samples)
 * is too short for the next read.  tolerant != 0: a window on which the reference itself raises
 * (LDD_FIELD_CRASH) is stepped over by 20 lines instead of ending the walk (range starts that are
 * not field-aligned: shards, chunks).  For every window the reference would have read it fills fields[i],
 * base[i] (plane index of the window start), winlen[i], readsample[i] (capture coordinates) and, for
 * LDD_FIELD_LOCATED fields, linelocs1/linebad rows.  A window that does not start on a peak of the
 * global chase gets its own peak list from `cb` (window-relative indices, host arrays that stay
 * valid until the next callback). */
typedef int (*ldd_window_peaks_fn)(void* ctx, long long plane_start, long long window_len,
                                   const long long** peaks, const double** vals, int* npeaks);
int ldd_field_chain(ldd_handle* h, const long long* gpeaks, const double* gvals, int ngpeaks,
                    long long plane_len, long long plane_origin, long long ncap, long long readlen,
                    long long first_readsample, long long stop_readsample, int tolerant,
                    int max_fields, ldd_window_peaks_fn cb, void* ctx,
                    ldd_field* fields, long long* base, long long* winlen, long long* readsample_out,
                    double* linelocs1, unsigned char* linebad, int ll_stride, int* nfields_out);

/* Field.refine_linelocs_hsync, batched: field f works on the window that starts at plane index
 * base_dev[f] and is winlen_dev[f] long; line tables are [nfields][ll_stride]; status_dev[f] (caller
 * zeroed) gets bit 1 when the reference would have raised (field invalid). */
int ldd_refine_hsync(ldd_handle* h, const float* d05_dev, long long n, const long long* base_dev,
                     const long long* winlen_dev, const int* linecount_dev, int nfields, int ll_stride,
                     const double* linelocs1_dev, const unsigned char* linebad_dev, double* linelocs2_dev,
                     unsigned char* linebad_out_dev, int* status_dev, void* stream);

/* FieldNTSC.refine_linelocs_burst (one pass; the reference runs it twice): linelocs_in -> linelocs_out,
 * burstlevel_dev float32 [nfields][ll_stride].  status bit 2: a line's geometry was not covered. */
int ldd_refine_burst(ldd_handle* h, const float* burst_dev, long long n, const long long* base_dev,
                     const int* linecount_dev, int nfields, int ll_stride, const double* linelocs_in_dev,
                     double* linelocs_out_dev, float* burstlevel_dev, int* status_dev, void* stream);

/* FieldPAL.refine_linelocs_pilot.  status bit 3: window outside the plane / too many crossings. */
int ldd_refine_pilot(ldd_handle* h, const float* demod_dev, const float* d05_dev, long long n,
                     const long long* base_dev, const int* linecount_dev, int nfields, int ll_stride,
                     const double* linelocs_in_dev, double* linelocs_out_dev, int* status_dev, void* stream);

/* ---- kernel (5): per-line TBC resampling.  Field.downscale + lddutils.scale (not-a-knot cubic
 * spline per line) + the uint16 quantisation of FieldNTSC/FieldPAL.downscale(final=True)
 * (lddecode_core.py:789-812, 1023-1035, 1135-1159; lddutils.py:83-97), batched over fields.
 *   plane_dev[n]           float32 plane; plane value + plane_add = Hz (ire0 for LDD_P_DEMOD, 0 for LDD_P_BURST)
 *   base_dev               [nfields] plane index of each field window's sample 0 (NULL: all 0)
 *   linelocs_dev           [nfields][ll_stride] float64 line positions relative to the window; output line k of
 *                          a field spans linelocs[lineoffset+k] .. linelocs[lineoffset+k+1]
 *   lineloc_add            constant added to every line position first (FieldNTSC.apply_offsets,
 *                          lddecode_core.py:1161-1162, 1186); 0 otherwise
 *   linecount_dev          [nfields] lines to produce per field (<= max_linecount)
 *   mode 0                 out_dev = float64 Hz, exactly what Field.downscale returns (dsout)
 *   mode 1                 out_dev = uint16 TBC samples; when burstlevel_dev (float32 [nfields][ll_stride]) is
 *                          given the NTSC burst markers are written into samples 0/1 of lines 1..linecount-2
 *   out_stride             elements between consecutive fields in out_dev
 *   status_dev             [nfields] int, caller-zeroed; bit 0 (1) set when a line was not resampled, together with
 *                          LDD_ST_LINE_LONG (32) when the only reason was a line longer than the staging window of this
 *                          pass (1.25 x nominal; lddutils.scale takes any length: ldd_tbc_long_lines does those lines) or
 *                          LDD_ST_LINE_BAD (64) when its window leaves the plane or is degenerate (the reference raises
 *                          and marks the field invalid) */
#define LDD_ST_LINE_LONG 32
#define LDD_ST_LINE_BAD 64
int ldd_tbc_fields(ldd_handle* h, const float* plane_dev, long long n, double plane_add, const long long* base_dev,
                   const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                   int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                   void* out_dev, long long out_stride, const float* burstlevel_dev, double colorlevel,
                   int* status_dev, void* stream);

/* The same with an explicit output layout: field f's line 0 starts at element out_off_dev[f] of out_dev (NULL:
 * f * out_stride) and consecutive lines are line_stride elements apart (0: outwidth).  With line_stride = 2 * outwidth and
 * the two fields of a frame at offsets 0 and outwidth this writes interleaved frames directly, i.e. Framer.formatoutput
 * (lddecode_core.py:1238-1252) without a second pass over the pictures. */
int ldd_tbc_fields_ex(ldd_handle* h, const float* plane_dev, long long n, double plane_add, const long long* base_dev,
                      const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                      int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                      void* out_dev, long long out_stride, const long long* out_off_dev, long long line_stride,
                      const float* burstlevel_dev, double colorlevel, int* status_dev, void* stream);

/* Second pass for the rare lines the pass above left out as too long for its staging window (status 1 | LDD_ST_LINE_LONG;
 * they come from fields whose line location partly failed -- the reference resamples whatever span it is given,
 * lddutils.py:83-97): same arguments, exact float64 kernel with room for spans up to 4032 samples, only those lines are
 * touched.  status_dev: a zeroed array of its own; a field that comes back 0 is complete. */
int ldd_tbc_long_lines(ldd_handle* h, const float* plane_dev, long long n, double plane_add, const long long* base_dev,
                       const double* linelocs_dev, int ll_stride, const int* linecount_dev, int nfields,
                       int max_linecount, int lineoffset, double lineloc_add, int outwidth, int wow, int mode,
                       void* out_dev, long long out_stride, const long long* out_off_dev, long long line_stride,
                       const float* burstlevel_dev, double colorlevel, int* status_dev, void* stream);

/* ---- VBI: Field.decodephillipscode (lddecode_core.py:814-834), batched.  For field f and code line lines[i]
 * (SysParams['philips_codelines']) codes_dev[4 f + i] receives the 24-bit Philips code (first cell = bit 23, i.e. the six
 * nibbles of the reference's `linecode`) or -1 where the reference returns None.  linelocs_dev is Field.linelocs at that
 * point of Field.__init__ (linelocs2); base/winlen as in ldd_refine_hsync (NULL: the whole plane is one window). */
int ldd_vbi_decode(ldd_handle* h, const float* demod_dev, long long n, const long long* base_dev,
                   const long long* winlen_dev, const double* linelocs_dev, int ll_stride, int nfields,
                   const int* lines, int nlines, int* codes_dev, void* stream);

/* ---- peer memory for the multi-GPU gather (one process per GPU, NVLink / NVSwitch).  The gather of per-field outputs to
 * the root rank needs no collective kernel on the root: the root allocates its receive buffer with
 * ldd_peer_alloc and publishes the 64-byte handle, every other rank maps it (ldd_peer_open) and passes addresses inside it
 * to ldd_pipe_finish as the TBC kernel's destination -- fields are stored straight into the root's HBM.  ldd_peer_signal
 * (a stream-ordered flag write behind a system-wide fence) and ldd_peer_wait (a kernel that spins until n flags, `stride`
 * ints apart, have reached `value`) order producers and consumer; ldd_peer_read is a blocking device-to-host copy.
 * ldd_peer_copy is the alternative to storing through the mapping from a kernel: one stream-ordered DMA transfer
 * (copy engine over NVLink, no SMs on either side) from a local buffer into the mapped one. */
int ldd_peer_alloc(size_t nbytes, void** dev_ptr, unsigned char* handle64);
int ldd_peer_open(const unsigned char* handle64, void** dev_ptr);
int ldd_peer_close(void* dev_ptr);
int ldd_peer_free(void* dev_ptr);
int ldd_peer_read(void* host_dst, const void* dev_src, size_t nbytes);
int ldd_peer_copy(void* dst_dev, const void* src_dev, size_t nbytes, void* stream);
int ldd_peer_signal(int* flag_dev, int value, void* stream);
int ldd_peer_wait(const int* flags_dev, int n, int stride, int value, void* stream);

/* ---- whole-range pipeline: Framer.readfield's loop (lddecode_core.py:1194-1223) over a range of a capture in two calls.
 * All buffers are the caller's.  One ldd_pipe per plane workspace; two of them software-pipeline a stream of ranges
 * (launch k+1, then finish k).  A pipe is used from one thread. */
typedef struct ldd_pipe ldd_pipe;

typedef struct ldd_pipe_bufs {
    /* device */
    void* planes[5];            /* LDD_P_*: plane_cap float32 each, float64 for LDD_P_SYNC; [LDD_P_PILOT] PAL only */
    long long plane_cap;
    double* audio1_l;           /* phase-1 audio, audio1_cap each; NULL: audio off */
    double* audio1_r;
    long long audio1_cap;
    double* audio2_l;           /* phase-2 audio, audio1_cap / 4 each */
    double* audio2_r;
    long long* peaks;           /* peak_cap */
    double* peak_vals;          /* peak_cap */
    int* peak_count;            /* 2 */
    int peak_cap;
    unsigned char* field_tables;    /* tables_bytes >= device_bytes of ldd_pipe_table_bytes */
    long long tables_bytes;
    /* page-locked host */
    long long* h_peaks;         /* peak_cap */
    double* h_peak_vals;        /* peak_cap */
    int* h_peak_count;          /* 2 */
    unsigned char* h_tables;    /* h_tables_bytes >= upload_bytes of ldd_pipe_table_bytes (field tables + PCM tables) */
    long long h_tables_bytes;
    double* h_prefix;           /* prefix_cap samples of demod_sync for a window the capture-wide chase does not decide */
    long long prefix_cap;       /* >= 40 * linelen */
} ldd_pipe_bufs;

typedef struct ldd_pipe_result {
    int nwindows;               /* windows the walk read (every readfield iteration of the reference) */
    int nowned;                 /* of those, read position in [r0, r1) */
    int nlocated;               /* owned windows that are fields (LDD_FIELD_LOCATED): refined and resampled, in order */
    int npeaks;                 /* peaks of the range-wide chase */
    int nframes;                /* frame mode: complete frames written */
    int prefix_windows;         /* windows that needed their own chase over a prefix (copied to the host) */
    int ll_stride;
    long long plane_origin, plane_len, walk_start;
    long long audio1_len, audio2_len;
    double lineloc_add;         /* FieldNTSC.apply_offsets term of the final line positions (0 for PAL) */
    /* host arrays owned by the pipe, valid until its next finish; indexed by window */
    const ldd_field* fields;
    const long long* base;
    const long long* winlen;
    const long long* readsample;
    const double* linelocs1;    /* [nwindows][ll_stride] */
    const unsigned char* linebad;
    const int* owned;           /* [nowned] window index */
    const int* located;         /* [nlocated] index into owned[] */
    const int* frame_of;        /* [nlocated] frame mode: frame index of the field, -1 = before the first frame */
    const long long* gpeaks;    /* the range-wide peak list (page-locked buffers of ldd_pipe_bufs) */
    const double* gvals;
    /* device tables inside ldd_pipe_bufs.field_tables, row k <-> located[k] */
    void* d_base; void* d_winlen; void* d_linecount; void* d_linelocs1;
    void* d_linelocs2; void* d_linebad2; void* d_linelocs3; void* d_linelocs4; void* d_burstlevel;
    void* d_final;              /* the line table the TBC used (before lineloc_add): linelocs4 NTSC, pilot-refined PAL */
    void* d_vbi;                /* int [nlocated][4]: ldd_vbi_decode codes of the three code lines */
} ldd_pipe_result;

int ldd_pipe_table_bytes(int max_fields, long long* upload_bytes, long long* device_bytes);
/* field_samples: int(freq_hz / FPS / 2), the nominal field length (range planning). */
int ldd_pipe_create(ldd_handle* h, const ldd_pipe_bufs* bufs, int max_fields, long long field_samples, ldd_pipe** out);
void ldd_pipe_destroy(ldd_pipe* p);
/* Stage 1, asynchronous on `stream`.  rf_dev holds capture samples [rf_base, rf_base + rf_len) in format fmt; ncap_total is the
 * length of the whole capture (the reference stops when a read would pass its end); the range owns the fields whose read
 * position is in [r0, r1).  readlen = Framer.readlen.  mtf_level is applied when LDD_F_MTF is set. */
int ldd_pipe_launch(ldd_pipe* p, const void* rf_dev, int fmt, long long rf_base, long long rf_len, long long ncap_total,
                    long long r0, long long r1, long long readlen, double mtf_level, int audio_phase2, void* stream);
/* Stage 2: blocks until the peak list of stage 1 is in host memory (not until the stream is idle), walks, and enqueues
 * refinement + VBI + TBC on refine_stream; main_stream (the stream of stage 1) is ordered behind them.  pic_dev receives
 * the uint16 fields: frame_mode 0 -> field k at k * pic_stride; frame_mode 1 -> interleaved frames of pic_stride
 * elements each (fields paired by parity as Framer.readframe does for CLV; fields before the first frame land in slot
 * pic_cap - 1).  pic_cap = fields / frames pic_dev can hold; status_dev [max_fields] gets the per-field error bits. */
int ldd_pipe_finish(ldd_pipe* p, double colorlevel, double colorphase, int frame_mode, void* pic_dev, long long pic_stride,
                    long long pic_cap, int* status_dev, void* refine_stream, void* main_stream, ldd_pipe_result* out);
/* ldd_tbc_long_lines over the fields of the last ldd_pipe_finish (same pictures, same layout): for ranges whose status came
 * back with 1 | LDD_ST_LINE_LONG and without LDD_ST_LINE_BAD.  status_dev: zeroed [nlocated] of its own. */
int ldd_pipe_long_lines(ldd_pipe* p, int* status_dev, void* stream);

/* Optional stage 3, asynchronous on `stream` (the main stream of stages 1 and 2): the 48 kHz PCM of the located fields of
 * the last ldd_pipe_finish, i.e. downscale_audio (lddecode_core.py:431-484) per field as Field.downscale(audio=True) calls
 * it (:809-810) on the final line positions, from the range's phase-2 audio (ldd_pipe_launch with audio_phase2 = 1).
 * The reference resamples each field from the audio of that field's own read window; its sample idx there lies at
 * (window start / decimation + idx) of the range's audio, in general between two samples, and is read by four-point
 * Lagrange interpolation (a window that starts on the range's audio grid gets the reference's very sample).  The time
 * offsets are chained on the host:
 *   LDD_PCM_CHAIN_FIELDS  every field starts where the previous one ended (audio_next_offset -> audio_offset);
 *   LDD_PCM_CHAIN_FRAMER  as Framer.readframe does for CLV discs (:1203, 1260-1289): all fields of one readframe call use the
 *                         offset the call started with, the field that closes the frame hands its audio_next_offset on,
 *                         and fields ahead of the first frame are dropped while bit 1 of *frame_state is set.
 * *audio_offset (seconds) and *frame_state (bit 0: a frame is open, bit 1: Framer's `firstframe`) are read and updated,
 * so consecutive ranges of one capture continue each other; start with 0.0 and 2.  out_dev receives interleaved L/R
 * int16, field k at out_off[k] .. out_off[k+1] (host array of nlocated + 1 entries, filled before the call returns;
 * an empty span = dropped field); status_dev [nlocated] gets bit 4 (16) where an index left the line table or the audio.
 * scale = the reference's `scale` argument (64), freq_hz its `freq` (48000). */
#define LDD_PCM_CHAIN_FIELDS 0
#define LDD_PCM_CHAIN_FRAMER 1
/* HOST function: the offset chain of ldd_pipe_pcm by itself, over the line counts and parities of nfields consecutive
 * fields -- what a rank of a sharded decode runs over the fields of the ranks before it to learn the state its own range
 * starts with (audio_next_offset is a prefix over per-field line counts).  nout (may be NULL)
 * receives the stereo samples every field contributes. */
int ldd_pcm_chain(int system, double freq_hz, double line_period_us, int chain, int nfields, const int* linecount,
                  const int* istop, double* audio_offset, int* frame_state, int* nout);
int ldd_pipe_pcm(ldd_pipe* p, double freq_hz, double scale, double line_period_us, double audio_lfreq, double audio_rfreq,
                 int chain, double* audio_offset, int* frame_state, short* out_dev, long long out_cap, long long* out_off,
                 int* status_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif
