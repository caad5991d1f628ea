// Whole-range decode in two calls: the steady state of a streaming / sharded decode without Python in it.
//
//   ldd_pipe_launch   plans the range's block grid (fixed global grid, halos), sets the MTF level, enqueues the fused
//                     demodulation, the sync-peak chase, the peak list's copy to page-locked memory and the second
//                     audio stage.  Asynchronous.
//   ldd_pipe_finish   waits for the peak list only, walks the fields on the host (Framer.readfield's walk,
//                     lddecode_core.py:1194-1223, via ldd_field_chain), fills the per-field tables in page-locked
//                     memory and enqueues upload, hsync / burst | pilot refinement, VBI decode and TBC for all located
//                     fields from preallocated device tables.  Returns while those kernels run.
//   ldd_pipe_pcm      (optional, after finish) 48 kHz PCM of the located fields from the range's phase-2 audio, with the
//                     time offsets chained the way Framer.readframe chains them (lddecode_core.py:1203, 1283-1289).
//
// The reference does this per field in Python (Framer.readfield -> RFDecode.demod -> FieldNTSC/FieldPAL.__init__);
// here the host's share of a step is the walk (~10 us per field) plus a dozen launches.  All buffers are the
// caller's (ldd_pipe_bufs); the pipe owns only two events, a side stream and host-side tables.
#include "ldd_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

using namespace ldd;

namespace {

constexpr int LL = 320;       // line-table stride (>= linecount + 4; PAL: 317)

struct TableLayout {
    // upload part (host -> device in one copy)
    size_t o_base, o_winlen, o_l1, o_linecount, o_bad, o_outoff, upload_bytes;
    // device-only part
    size_t o_l2, o_bad2, o_l3, o_l4, o_bl, o_vbi, o_status_unused;
    // tables of ldd_pipe_pcm: on the device at o_pcm, in the page-locked staging behind the upload part; r_* are offsets
    // inside that region
    size_t o_pcm, r_t0, r_t1, r_fbase, r_off, r_nout, pcm_bytes, total_bytes;
};

TableLayout layout(int F) {
    TableLayout t;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 63) & ~(size_t)63; return at; };
    t.o_base = take((size_t)F * 8);
    t.o_winlen = take((size_t)F * 8);
    t.o_l1 = take((size_t)F * LL * 8);
    t.o_linecount = take((size_t)F * 4);
    t.o_bad = take((size_t)F * LL);
    t.o_outoff = take((size_t)F * 8);
    t.upload_bytes = o;
    t.o_l2 = take((size_t)F * LL * 8);
    t.o_bad2 = take((size_t)F * LL);
    t.o_l3 = take((size_t)F * LL * 8);
    t.o_l4 = take((size_t)F * LL * 8);
    t.o_bl = take((size_t)F * LL * 4);
    t.o_vbi = take((size_t)F * 4 * 4);
    t.o_status_unused = o;
    t.o_pcm = o;
    t.r_t0 = take((size_t)F * 8) - t.o_pcm;
    t.r_t1 = take((size_t)F * 8) - t.o_pcm;
    t.r_fbase = take((size_t)F * 8) - t.o_pcm;
    t.r_off = take((size_t)F * 8) - t.o_pcm;
    t.r_nout = take((size_t)F * 4) - t.o_pcm;
    t.pcm_bytes = o - t.o_pcm;
    t.total_bytes = o;
    return t;
}

}  // namespace

struct ldd_pipe {
    ldd_handle* h;
    ldd_pipe_bufs b;
    int max_fields;
    long long field_samples;
    TableLayout lay;
    cudaEvent_t ev_peaks = nullptr;     // recorded behind the peak list's copy to page-locked memory
    cudaEvent_t ev_upload = nullptr;    // the table upload from h_tables has executed
    cudaEvent_t ev_done = nullptr;      // refine + TBC of the last finish
    cudaEvent_t ev_pcm = nullptr;       // the PCM tables' upload has executed
    bool pcm_pending = false;
    // what ldd_pipe_pcm needs of the last finish
    bool finished = false;
    const void* fin_final = nullptr;
    const void* fin_linecount = nullptr;
    double fin_lineloc_add = 0.0;
    // the TBC launch of the last finish, for ldd_pipe_long_lines
    struct { const float* plane; long long plen; const long long* base; const double* ll; const int* lc; int n, maxlc, lineoffset;
             double add; void* pic; long long pic_stride; const long long* off; long long line_stride; const float* bl;
             double colorlevel; } tbc = {};
    cudaStream_t side = nullptr;        // prefix copies of off-chain windows
    bool upload_pending = false;
    long long audio1_len = 0, audio2_len = 0;
    // plan of the launched range
    bool launched = false;
    long long r0 = 0, r1 = 0, ncap_total = 0, readlen = 0, plane_origin = 0, plane_len = 0, walk_start = 0;
    // host tables of the walk
    std::vector<ldd_field> fields;
    std::vector<long long> base, winlen, readsample;
    std::vector<double> linelocs1;
    std::vector<unsigned char> linebad;
    std::vector<int> owned, located, frame_of;
    // an off-chain window's own peak list
    std::vector<long long> wpk, spk;
    std::vector<double> wvl, svl;
    int prefix_windows = 0;             // how many windows of the last walk needed their own chase
};

namespace {

// The time-offset chain of the fields' PCM.  One field: np.arange(timeoffset, frametime + soundgap, soundgap) has
// ceil((stop - start) / step) values start + i * ((start + step) - start); the field produces one sample less and hands on
// arange[-1] - frametime (lddecode_core.py:432-438, 482).  LDD_PCM_CHAIN_FRAMER is Framer.readframe (:1203, 1260-1289) with
// the CLV pairing: every field of a call is built with the offset the call started with; a field whose parity is
// `topfirst` opens the frame, the next one closes it and the closing field's offset is carried on; fields ahead of the
// very first frame are not written.
struct PcmChain {
    double off;
    bool open, first;
    // returns the stereo samples the field contributes; *t0 / *t1 = the first two values of its arange
    int step(int linecount, int istop, int topfirst, double freq_hz, double line_period_us, int chain, double* t0, double* t1) {
        const double soundgap = 1.0 / freq_hz;
        const double frametime = (line_period_us * (double)linecount) / 1000000.0;
        const double stop = frametime + soundgap;
        const double lenf = std::ceil((stop - off) / soundgap);
        const long long len = lenf > 0 ? (long long)lenf : 0;
        const double second = off + soundgap, delta = second - off;
        const double last = len <= 1 ? off : (len == 2 ? second : off + (double)(len - 1) * delta);
        const double next = len >= 1 ? last - frametime : off;
        bool include = true, closes = false;
        if (chain == LDD_PCM_CHAIN_FRAMER) {
            if (istop == topfirst) open = true;
            else if (open) closes = true;
            include = open || closes || !first;
        }
        *t0 = off; *t1 = second;
        const int nout = include && len > 1 ? (int)(len - 1) : 0;
        if (chain == LDD_PCM_CHAIN_FIELDS) off = next;
        else if (closes) { off = next; open = false; first = false; }
        return nout;
    }
};

int pfail(ldd_pipe* p, int code, const char* msg) {
    if (p && p->h) p->h->err = msg;
    return code;
}

// Peak list of a window that the global chase does not decide: chase a short prefix on the samples themselves (copied
// over on the side stream), splice into the global list at the first shared peak, cut at the reference's loop bound.
int pipe_window_peaks(void* ctx, long long b, long long wl, const long long** peaks, const double** vals, int* npeaks) {
    ldd_pipe* p = (ldd_pipe*)ctx;
    ldd_handle* h = p->h;
    const int L = h->cfg.linelen;
    const long long half = L / 2, skip = (long long)(L * .4);
    const long long* gpk = p->b.h_peaks;
    const double* gvl = p->b.h_peak_vals;
    const int ng = std::min(p->b.h_peak_count[0], p->b.peak_cap);
    const double* sync = (const double*)p->b.planes[LDD_P_SYNC];
    ++p->prefix_windows;
    long long npre = std::min<long long>(wl, std::min<long long>(40LL * L, p->b.prefix_cap));
    int nsp = 0;
    if (npre > 2LL * L) {
        if (cudaMemcpyAsync(p->b.h_prefix, sync + b, (size_t)npre * sizeof(double), cudaMemcpyDeviceToHost, p->side) != cudaSuccess ||
            cudaStreamSynchronize(p->side) != cudaSuccess)
            return pfail(p, LDD_ECUDA, "prefix copy failed");
        const int cap = (int)(npre / skip) + 8;
        p->spk.resize(cap); p->svl.resize(cap);
        int rc = ldd_sync_peaks_host(h, p->b.h_prefix, npre, 0, p->spk.data(), p->svl.data(), cap, &nsp);
        if (rc) return rc;
        nsp = std::min(nsp, cap);
    }
    int k = -1, m = -1;
    for (int i = 0; i < nsp; ++i) {
        const long long* it = std::lower_bound(gpk, gpk + ng, p->spk[i] + b);
        if (it != gpk + ng && *it == p->spk[i] + b) { k = i; m = (int)(it - gpk); break; }
    }
    p->wpk.clear(); p->wvl.clear();
    if (k >= 0) {
        for (int i = 0; i < k; ++i) { p->wpk.push_back(p->spk[i]); p->wvl.push_back(p->svl[i]); }
        for (int i = m; i < ng; ++i) { p->wpk.push_back(gpk[i] - b); p->wvl.push_back(gvl[i]); }
        // a peak belongs to the window's list iff the step that found it started below the loop bound
        const long long limit = wl - 2LL * L;
        size_t n = p->wpk.size();
        for (size_t i = 0; i < p->wpk.size(); ++i) {
            const long long i0 = i == 0 ? 0 : p->wpk[i - 1] + skip;
            const long long istep = i0 + ((p->wpk[i] - i0) / half) * half;
            if (istep >= limit) { n = i; break; }
        }
        p->wpk.resize(n); p->wvl.resize(n);
    } else {
        // never merged inside the prefix: chase the whole window on the device (side stream; the planes are complete)
        // (rare: the chase's workspace belongs to the handle, so nothing else may be in flight)
        const int cap = p->b.peak_cap;
        cudaDeviceSynchronize();
        int rc = ldd_sync_peaks(h, sync + b, wl, 0, p->b.peaks, p->b.peak_vals, cap, p->b.peak_count, p->side);
        if (rc) return rc;
        int cnt[2] = {0, 0};
        cudaMemcpyAsync(cnt, p->b.peak_count, sizeof cnt, cudaMemcpyDeviceToHost, p->side);
        cudaStreamSynchronize(p->side);
        const int n = std::min(cnt[0], cap);
        p->wpk.resize(n); p->wvl.resize(n);
        cudaMemcpyAsync(p->wpk.data(), p->b.peaks, (size_t)n * 8, cudaMemcpyDeviceToHost, p->side);
        cudaMemcpyAsync(p->wvl.data(), p->b.peak_vals, (size_t)n * 8, cudaMemcpyDeviceToHost, p->side);
        if (cudaStreamSynchronize(p->side) != cudaSuccess) return pfail(p, LDD_ECUDA, "window chase failed");
    }
    *peaks = p->wpk.data();
    *vals = p->wvl.data();
    *npeaks = (int)p->wpk.size();
    return LDD_OK;
}

}  // namespace

extern "C" {

int ldd_pipe_table_bytes(int max_fields, long long* upload_bytes, long long* device_bytes) {
    if (max_fields < 1) return LDD_EINVAL;
    TableLayout t = layout(max_fields);
    if (upload_bytes) *upload_bytes = (long long)(t.upload_bytes + t.pcm_bytes);
    if (device_bytes) *device_bytes = (long long)t.total_bytes;
    return LDD_OK;
}

int ldd_pipe_create(ldd_handle* h, const ldd_pipe_bufs* bufs, int max_fields, long long field_samples, ldd_pipe** out) {
    if (!h || !bufs || !out || max_fields < 1 || field_samples < 1) return LDD_EINVAL;
    *out = nullptr;
    const bool pal = h->cfg.system == LDD_SYSTEM_PAL;
    for (int i = 0; i < (pal ? 5 : 4); ++i)
        if (!bufs->planes[i]) { h->err = "ldd_pipe_create: plane buffer missing"; return LDD_EINVAL; }
    if (!bufs->peaks || !bufs->peak_vals || !bufs->peak_count || !bufs->h_peaks || !bufs->h_peak_vals || !bufs->h_peak_count ||
        !bufs->field_tables || !bufs->h_tables || !bufs->h_prefix || bufs->peak_cap < 16) {
        h->err = "ldd_pipe_create: buffer missing";
        return LDD_EINVAL;
    }
    TableLayout t = layout(max_fields);
    if ((size_t)bufs->tables_bytes < t.total_bytes || (size_t)bufs->h_tables_bytes < t.upload_bytes + t.pcm_bytes) {
        h->err = "ldd_pipe_create: table buffers too small (ldd_pipe_table_bytes)";
        return LDD_ECAP;
    }
    ldd_pipe* p = new (std::nothrow) ldd_pipe();
    if (!p) return LDD_ENOMEM;
    p->h = h; p->b = *bufs; p->max_fields = max_fields; p->field_samples = field_samples; p->lay = t;
    p->fields.resize(max_fields); p->base.resize(max_fields); p->winlen.resize(max_fields); p->readsample.resize(max_fields);
    p->linelocs1.resize((size_t)max_fields * LL); p->linebad.resize((size_t)max_fields * LL);
    p->owned.reserve(max_fields); p->located.reserve(max_fields); p->frame_of.reserve(max_fields);
#ifndef LDD_EMU
    if (cudaEventCreateWithFlags(&p->ev_peaks, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_upload, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_pcm, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithFlags(&p->side, cudaStreamNonBlocking) != cudaSuccess) {
        h->err = "ldd_pipe_create: event / stream creation failed";
        delete p;
        return LDD_ECUDA;
    }
#endif
    *out = p;
    return LDD_OK;
}

void ldd_pipe_destroy(ldd_pipe* p) {
    if (!p) return;
#ifndef LDD_EMU
    if (p->ev_peaks) cudaEventDestroy(p->ev_peaks);
    if (p->ev_upload) cudaEventDestroy(p->ev_upload);
    if (p->ev_done) cudaEventDestroy(p->ev_done);
    if (p->ev_pcm) cudaEventDestroy(p->ev_pcm);
    if (p->side) cudaStreamDestroy(p->side);
#endif
    delete p;
}

int ldd_pipe_launch(ldd_pipe* p, const void* rf_dev, int fmt, long long rf_base, long long rf_len, long long ncap_total,
                    long long r0, long long r1, long long readlen, double mtf_level, int audio_phase2, void* stream) {
    if (!p || !rf_dev || r1 <= r0 || readlen < 1) return LDD_EINVAL;
    ldd_handle* h = p->h;
    const ldd_config& c = h->cfg;
    const long long N = c.blocklen, bc = c.blockcut, S = N - bc - c.blockcut_end;
    cudaStream_t st = (cudaStream_t)stream;
    // block grid of the range owning read positions [r0, r1): it starts walking 1.6 fields early (by then its read
    // positions are the sequential walk's) and demodulates one read length past its end
    const long long walk_start = r0 <= 0 ? 0 : std::max<long long>(0, r0 - (long long)(1.6 * (double)p->field_samples));
    const long long first_block = (std::max<long long>(walk_start - bc, 0) / S) * S;
    const long long need_end = std::min(ncap_total, r1 + readlen + 2 * N + bc);
    long long nblocks = std::max<long long>(0, (need_end - first_block - N) / S + 1);
    const long long avail_end = std::min(rf_base + rf_len, ncap_total);
    while (nblocks > 0 && first_block + (nblocks - 1) * S + N > avail_end) --nblocks;
    if (first_block < rf_base) return pfail(p, LDD_EINVAL, "capture window does not cover the range's halo");
    // The range reaches the end of the capture and the capture does not end on the block grid: one more block, zeros past
    // the end.  Without it the planes stop up to one block stride short of the last window the reference still reads.
    bool pad_last = false;
    if (nblocks > 0 && need_end == ncap_total && avail_end == ncap_total && first_block + (nblocks - 1) * S + N < ncap_total &&
        (nblocks + 1) * S <= p->b.plane_cap) {
        ++nblocks;
        pad_last = true;
    }
    long long total = nblocks * S;
    // a padded last block: the planes end where the capture ends (what lies beyond is the transform of zeros)
    if (pad_last && total > ncap_total - first_block - bc) total = ncap_total - first_block - bc;
    if (total > p->b.plane_cap) return pfail(p, LDD_ECAP, "plane buffers too small for this range");
    p->r0 = r0; p->r1 = r1; p->ncap_total = ncap_total; p->readlen = readlen;
    p->plane_origin = first_block; p->plane_len = total; p->walk_start = walk_start;
    int rc;
    if (h->have_filter[LDD_F_MTF]) {
        rc = ldd_set_mtf_level(h, mtf_level, stream);
        if (rc) return rc;
    }
    const bool audio = h->A > 0 && p->b.audio1_l && p->b.audio1_r;
    long long alen = 0;
    if (audio) {
        alen = total / (N / h->A);
        if (alen > p->b.audio1_cap) return pfail(p, LDD_ECAP, "audio buffers too small for this range");
    }
    if (nblocks) {
        rc = (pad_last ? demod_blocks_padded : ldd_demod_blocks)(h, rf_dev, fmt, rf_base, rf_len, first_block, nblocks, total, p->b.planes,
                                                                 audio ? p->b.audio1_l : nullptr, audio ? p->b.audio1_r : nullptr, alen, stream);
        if (rc) return rc;
    }
    // sync-peak chase over the whole plane; its list is the only device -> host hop in the middle of the path.  It goes
    // to page-locked memory through a kernel, not the copy engine (there it would queue behind a field download).
    rc = ldd_sync_peaks(h, (const double*)p->b.planes[LDD_P_SYNC], total, 0, p->b.peaks, p->b.peak_vals, p->b.peak_cap, p->b.peak_count, stream);
    if (rc) return rc;
    rc = ldd_peaks_to_host(p->b.peaks, p->b.peak_vals, p->b.peak_count, p->b.peak_cap, p->b.h_peaks, p->b.h_peak_vals, p->b.h_peak_count, stream);
    if (rc) return rc;
    if (cudaEventRecord(p->ev_peaks, st) != cudaSuccess) return pfail(p, LDD_ECUDA, "cudaEventRecord");
    // the second audio stage does not depend on the walk: it runs behind the chase, under the host walk
    p->audio2_len = 0;
    if (audio && audio_phase2 && alen > N && p->b.audio2_l && p->b.audio2_r) {
        rc = ldd_audio_phase2(h, p->b.audio1_l, p->b.audio1_r, alen, p->b.audio2_l, p->b.audio2_r, stream);
        if (rc) return rc;
        p->audio2_len = alen / 4;
    }
    p->audio1_len = alen;
    p->launched = true;
    return LDD_OK;
}

int ldd_pipe_finish(ldd_pipe* p, double colorlevel, double colorphase, int frame_mode, void* pic_dev, long long pic_stride,
                    long long pic_cap, int* status_dev, void* refine_stream, void* main_stream, ldd_pipe_result* out) {
    if (!p || !out || !pic_dev || !status_dev) return LDD_EINVAL;
    if (!p->launched) return pfail(p, LDD_EINVAL, "ldd_pipe_finish without ldd_pipe_launch");
    p->launched = false;
    p->finished = false;
    ldd_handle* h = p->h;
    const ldd_config& c = h->cfg;
    const bool pal = c.system == LDD_SYSTEM_PAL;
    const int W = c.outlinelen;
    memset(out, 0, sizeof *out);
    if (cudaEventSynchronize(p->ev_peaks) != cudaSuccess) return pfail(p, LDD_ECUDA, "peak event");
    int ng = p->b.h_peak_count[0];
    if (ng > p->b.peak_cap) return pfail(p, LDD_ECAP, "peak list truncated: peak_cap too small");
    // ---- host walk
    int nf = 0;
    p->prefix_windows = 0;
    int rc = ldd_field_chain(h, p->b.h_peaks, p->b.h_peak_vals, ng, p->plane_len, p->plane_origin, p->ncap_total, p->readlen,
                             p->walk_start, p->r1, p->r0 > 0 ? 1 : 0, p->max_fields, pipe_window_peaks, p, p->fields.data(),
                             p->base.data(), p->winlen.data(), p->readsample.data(), p->linelocs1.data(), p->linebad.data(), LL, &nf);
    if (rc) return rc;
    p->owned.clear(); p->located.clear(); p->frame_of.clear();
    for (int i = 0; i < nf; ++i)
        if (p->readsample[i] >= p->r0 && p->readsample[i] < p->r1) {
            if (p->fields[i].stage == LDD_FIELD_LOCATED) p->located.push_back((int)p->owned.size());
            p->owned.push_back(i);
        }
    const int n = (int)p->located.size();
    out->nwindows = nf; out->nowned = (int)p->owned.size(); out->nlocated = n; out->npeaks = ng;
    out->prefix_windows = p->prefix_windows;
    out->plane_origin = p->plane_origin; out->plane_len = p->plane_len; out->walk_start = p->walk_start;
    out->fields = p->fields.data(); out->base = p->base.data(); out->winlen = p->winlen.data(); out->readsample = p->readsample.data();
    out->linelocs1 = p->linelocs1.data(); out->linebad = p->linebad.data(); out->ll_stride = LL;
    out->owned = p->owned.data(); out->located = p->located.data();
    out->gpeaks = p->b.h_peaks; out->gvals = p->b.h_peak_vals;
    out->audio1_len = p->audio1_len; out->audio2_len = p->audio2_len;
    const TableLayout& t = p->lay;
    unsigned char* dt = p->b.field_tables;
    out->d_base = dt + t.o_base; out->d_winlen = dt + t.o_winlen; out->d_linelocs1 = dt + t.o_l1; out->d_linecount = dt + t.o_linecount;
    out->d_linelocs2 = dt + t.o_l2; out->d_linebad2 = dt + t.o_bad2; out->d_linelocs3 = dt + t.o_l3; out->d_linelocs4 = dt + t.o_l4;
    out->d_burstlevel = dt + t.o_bl; out->d_vbi = dt + t.o_vbi;
    out->d_final = pal ? out->d_linelocs3 : out->d_linelocs4;
    const double shift33 = colorphase * (3.14159265358979323846 / 180.0);
    out->lineloc_add = pal ? 0.0 : (shift33 - 8) * ((c.freq_hz / 1e6) / (4.0 * 315.0 / 88.0));
    out->frame_of = nullptr;
    p->fin_final = out->d_final; p->fin_linecount = out->d_linecount; p->fin_lineloc_add = out->lineloc_add;
    if (n == 0) { p->finished = true; return LDD_OK; }
    // ---- per-field tables of the located fields -> page-locked staging
    if (p->upload_pending) { cudaEventSynchronize(p->ev_upload); p->upload_pending = false; }
    unsigned char* ht = p->b.h_tables;
    long long* hb = (long long*)(ht + t.o_base);
    long long* hw = (long long*)(ht + t.o_winlen);
    double* hl1 = (double*)(ht + t.o_l1);
    int* hlc = (int*)(ht + t.o_linecount);
    unsigned char* hbad = ht + t.o_bad;
    long long* hoff = (long long*)(ht + t.o_outoff);
    int maxlc = 0;
    // frame mode (Framer.readframe's pairing for CLV / parity, lddecode_core.py:1272-1281, and formatoutput, :1238-1252):
    // a frame starts with a field whose istop == topfirst; the top field's line i is frame line 2i, the other field's 2i+1
    const int topfirst = pal ? 0 : 1;
    int nframes = 0, open_frame = -1;
    p->frame_of.assign(n, -1);
    for (int k = 0; k < n; ++k) {
        const int w = p->owned[p->located[k]];
        const ldd_field& f = p->fields[w];
        hb[k] = p->base[w]; hw[k] = p->winlen[w]; hlc[k] = f.linecount;
        memcpy(hl1 + (size_t)k * LL, &p->linelocs1[(size_t)w * LL], LL * sizeof(double));
        memcpy(hbad + (size_t)k * LL, &p->linebad[(size_t)w * LL], LL);
        maxlc = std::max(maxlc, f.linecount);
        if (frame_mode) {
            long long off = -1;
            if (f.istop == topfirst) { open_frame = nframes++; }
            if (open_frame >= 0) {
                off = (long long)open_frame * pic_stride + (f.istop ? 0 : W);
                p->frame_of[k] = open_frame;
                if (f.istop != topfirst) open_frame = -1;
            }
            // fields before the first frame start go to the spare slot behind the last frame the buffer can hold
            hoff[k] = off >= 0 ? off : (pic_cap - 1) * pic_stride + (f.istop ? 0 : W);
        } else {
            hoff[k] = (long long)k * pic_stride;
        }
    }
    if ((frame_mode ? nframes + 1 : n) > pic_cap) return pfail(p, LDD_ECAP, "picture buffer too small");
    out->nframes = open_frame >= 0 ? nframes - 1 : nframes;       // complete frames; a trailing top field alone has frame_of >= nframes
    out->frame_of = p->frame_of.data();
    // ---- device work, on refine_stream (ordered behind the demodulation through the peak event)
    cudaStream_t rs = (cudaStream_t)refine_stream, ms = (cudaStream_t)main_stream;
    if (rs != ms) cudaStreamWaitEvent(rs, p->ev_peaks, 0);
    rc = ldd_copy_small(dt, ht, t.upload_bytes, rs);
    if (rc) return pfail(p, rc, "table upload failed");
    cudaEventRecord(p->ev_upload, rs);
    p->upload_pending = true;
    cudaMemsetAsync(status_dev, 0, (size_t)n * sizeof(int), rs);
    const long long plen = p->plane_len;
    const long long* d_base = (const long long*)out->d_base;
    const long long* d_win = (const long long*)out->d_winlen;
    const int* d_lc = (const int*)out->d_linecount;
    const float* pl_demod = (const float*)p->b.planes[LDD_P_DEMOD];
    const float* pl_d05 = (const float*)p->b.planes[LDD_P_DEMOD05];
    rc = ldd_refine_hsync(h, pl_d05, plen, d_base, d_win, d_lc, n, LL, (const double*)out->d_linelocs1, dt + t.o_bad,
                          (double*)out->d_linelocs2, (unsigned char*)out->d_linebad2, status_dev, rs);
    if (rc) return rc;
    static const int lines_pal[3] = {19, 20, 21}, lines_ntsc[3] = {16, 17, 18};      // SysParams['philips_codelines']
    const int* lines = pal ? lines_pal : lines_ntsc;
    rc = ldd_vbi_decode(h, pl_demod, plen, d_base, d_win, (const double*)out->d_linelocs2, LL, n, lines, 3, (int*)out->d_vbi, rs);
    if (rc) return rc;
    const long long line_stride = frame_mode ? 2LL * W : W;
    if (!pal) {
        const float* pl_burst = (const float*)p->b.planes[LDD_P_BURST];
        rc = ldd_refine_burst(h, pl_burst, plen, d_base, d_lc, n, LL, (const double*)out->d_linelocs2, (double*)out->d_linelocs3,
                              (float*)out->d_burstlevel, status_dev, rs);
        if (rc) return rc;
        rc = ldd_refine_burst(h, pl_burst, plen, d_base, d_lc, n, LL, (const double*)out->d_linelocs3, (double*)out->d_linelocs4,
                              (float*)out->d_burstlevel, status_dev, rs);
        if (rc) return rc;
        rc = ldd_tbc_fields_ex(h, pl_demod, plen, c.ire0, d_base, (const double*)out->d_linelocs4, LL, d_lc, n, maxlc, 1, out->lineloc_add,
                               W, 1, 1, pic_dev, pic_stride, (const long long*)(dt + t.o_outoff), line_stride,
                               (const float*)out->d_burstlevel, colorlevel, status_dev, rs);
        p->tbc = {pl_demod, plen, d_base, (const double*)out->d_linelocs4, d_lc, n, maxlc, 1, out->lineloc_add, pic_dev, pic_stride,
                  (const long long*)(dt + t.o_outoff), line_stride, (const float*)out->d_burstlevel, colorlevel};
    } else {
        rc = ldd_refine_pilot(h, pl_demod, pl_d05, plen, d_base, d_lc, n, LL, (const double*)out->d_linelocs2, (double*)out->d_linelocs3,
                              status_dev, rs);
        if (rc) return rc;
        rc = ldd_tbc_fields_ex(h, pl_demod, plen, c.ire0, d_base, (const double*)out->d_linelocs3, LL, d_lc, n, maxlc, 3, 0.0, W, 1, 1,
                               pic_dev, pic_stride, (const long long*)(dt + t.o_outoff), line_stride, nullptr, colorlevel, status_dev, rs);
        p->tbc = {pl_demod, plen, d_base, (const double*)out->d_linelocs3, d_lc, n, maxlc, 3, 0.0, pic_dev, pic_stride,
                  (const long long*)(dt + t.o_outoff), line_stride, nullptr, colorlevel};
    }
    if (rc) return rc;
    if (rs != ms) {
        cudaEventRecord(p->ev_done, rs);
        cudaStreamWaitEvent(ms, p->ev_done, 0);
    }
    p->finished = true;
    return LDD_OK;
}

int ldd_pipe_long_lines(ldd_pipe* p, int* status_dev, void* stream) {
    if (!p || !status_dev) return LDD_EINVAL;
    if (!p->finished) return pfail(p, LDD_EINVAL, "ldd_pipe_long_lines without ldd_pipe_finish");
    if (p->located.empty()) return LDD_OK;
    const auto& a = p->tbc;
    return ldd_tbc_long_lines(p->h, a.plane, a.plen, p->h->cfg.ire0, a.base, a.ll, LL, a.lc, a.n, a.maxlc, a.lineoffset, a.add,
                              p->h->cfg.outlinelen, 1, 1, a.pic, a.pic_stride, a.off, a.line_stride, a.bl, a.colorlevel, status_dev, stream);
}

int ldd_pcm_chain(int system, double freq_hz, double line_period_us, int chain, int nfields, const int* linecount,
                  const int* istop, double* audio_offset, int* frame_state, int* nout) {
    if (!audio_offset || !frame_state || nfields < 0 || (nfields && (!linecount || !istop)) || freq_hz <= 0 || line_period_us <= 0 ||
        (chain != LDD_PCM_CHAIN_FIELDS && chain != LDD_PCM_CHAIN_FRAMER))
        return LDD_EINVAL;
    PcmChain ch{*audio_offset, (*frame_state & 1) != 0, (*frame_state & 2) != 0};
    const int topfirst = system == LDD_SYSTEM_PAL ? 0 : 1;
    for (int k = 0; k < nfields; ++k) {
        double t0, t1;
        const int n = ch.step(linecount[k], istop[k], topfirst, freq_hz, line_period_us, chain, &t0, &t1);
        if (nout) nout[k] = n;
    }
    *audio_offset = ch.off;
    *frame_state = (ch.open ? 1 : 0) | (ch.first ? 2 : 0);
    return LDD_OK;
}

int ldd_pipe_pcm(ldd_pipe* p, double freq_hz, double scale, double line_period_us, double audio_lfreq, double audio_rfreq,
                 int chain, double* audio_offset, int* frame_state, short* out_dev, long long out_cap, long long* out_off,
                 int* status_dev, void* stream) {
    if (!p || !audio_offset || !frame_state || !out_dev || !out_off || !status_dev || freq_hz <= 0 || scale <= 0 ||
        line_period_us <= 0 || (chain != LDD_PCM_CHAIN_FIELDS && chain != LDD_PCM_CHAIN_FRAMER))
        return LDD_EINVAL;
    if (!p->finished) return pfail(p, LDD_EINVAL, "ldd_pipe_pcm without ldd_pipe_finish");
    if (p->audio2_len <= 0 || !p->b.audio2_l || !p->b.audio2_r) return pfail(p, LDD_EINVAL, "ldd_pipe_pcm needs the range's phase-2 audio");
    ldd_handle* h = p->h;
    const ldd_config& c = h->cfg;
    const int n = (int)p->located.size();
    out_off[0] = 0;
    if (n == 0) return LDD_OK;
    const TableLayout& t = p->lay;
    if (p->pcm_pending) { cudaEventSynchronize(p->ev_pcm); p->pcm_pending = false; }
    unsigned char* hp = p->b.h_tables + t.upload_bytes;
    double* ht0 = (double*)(hp + t.r_t0);
    double* ht1 = (double*)(hp + t.r_t1);
    double* hfb = (double*)(hp + t.r_fbase);
    long long* hoff = (long long*)(hp + t.r_off);
    int* hn = (int*)(hp + t.r_nout);
    // plane samples per phase-2 audio sample: the first stage keeps every (N / A)-th, the second every 4th
    const double dec = 4.0 * (double)(c.blocklen / h->A);
    const int topfirst = c.system == LDD_SYSTEM_PAL ? 0 : 1;
    PcmChain ch{*audio_offset, (*frame_state & 1) != 0, (*frame_state & 2) != 0};
    int maxn = 0;
    for (int k = 0; k < n; ++k) {
        const int w = p->owned[p->located[k]];
        const ldd_field& f = p->fields[w];
        double t0, t1;
        hn[k] = ch.step(f.linecount, f.istop, topfirst, freq_hz, line_period_us, chain, &t0, &t1);
        ht0[k] = t0; ht1[k] = t1;
        hfb[k] = (double)p->base[w] / dec;
        hoff[k] = out_off[k];
        out_off[k + 1] = out_off[k] + 2LL * hn[k];
        maxn = std::max(maxn, hn[k]);
    }
    const double off = ch.off;
    const bool open = ch.open, first = ch.first;
    if (out_off[n] > out_cap) return pfail(p, LDD_ECAP, "PCM buffer too small");
    *audio_offset = off;
    *frame_state = (open ? 1 : 0) | (first ? 2 : 0);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* dp = p->b.field_tables + t.o_pcm;
    int rc = ldd_copy_small(dp, hp, t.pcm_bytes, st);
    if (rc) return pfail(p, rc, "PCM table upload failed");
    cudaEventRecord(p->ev_pcm, st);
    p->pcm_pending = true;
    return pcm_range_launch(h, p->b.audio2_l, p->b.audio2_r, p->audio2_len, (const double*)(dp + t.r_fbase), (const double*)p->fin_final,
                            LL, (const int*)p->fin_linecount, (const double*)(dp + t.r_t0), (const double*)(dp + t.r_t1),
                            (const int*)(dp + t.r_nout), (const long long*)(dp + t.r_off), n, maxn, p->fin_lineloc_add, scale,
                            line_period_us, audio_lfreq, audio_rfreq, out_dev, status_dev, st);
}

}  // extern "C"
