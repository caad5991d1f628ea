// Philips code (VBI) decode, Field.decodephillipscode (lddecode_core.py:814-834), batched: one warp per
// (field, code line).  The line's samples are staged in shared memory with coalesced loads, lane 0 then walks the
// 24 biphase cells exactly as the reference does (first crossing of 50 IRE within 12 us of line start + 2 us, then
// one crossing every ~2 us, the bit being the level half a microsecond before the crossing), and the gap check
// (1.85 .. 2.15 us) decides validity.  The interpretation of the three codes (processphilipscode, :836-884) is a
// handful of integer tests per field and stays with the caller.
#include "ldd_internal.h"

namespace ldd {

constexpr int VBI_STAGE = 2560;       // samples staged per line: (2 + 12 + 24 * 2.15 + 1) us at <= 38 MSPS

struct VbiParams {
    const float* demod;       // demod plane, relative to ire0
    long long n;
    const long long* base;    // [nfields] plane index of the field window's sample 0 (NULL: 0)
    const long long* winlen;  // [nfields] window length (NULL: the plane)
    const double* linelocs;   // [nfields][ll_stride]: Field.linelocs at the time of the decode (linelocs2)
    int ll_stride;
    int lines[4];             // SysParams['philips_codelines']
    int nlines;
    double freq;              // MHz
    double thr_rel;           // iretohz(50) - ire0
    int* codes;               // [nfields][4]: 24-bit code (first cell = bit 23) or -1 (None)
};

__global__ void __launch_bounds__(32) vbi_kernel(const VbiParams p) {
    __shared__ float s_d[VBI_STAGE];
    const int f = blockIdx.y, li = blockIdx.x, lane = threadIdx.x;
    const long long base = p.base ? p.base[f] : 0;
    long long len = p.winlen ? p.winlen[f] : p.n - base;
    if (base + len > p.n) len = p.n - base;
    const double linestart = p.linelocs[(size_t)f * p.ll_stride + p.lines[li]];
    const double fq = p.freq;
    const long long s0 = (long long)(linestart + 2.0 * fq) - (long long)(fq) - 2;     // staged from a little before the first scan
    for (int i = lane; i < VBI_STAGE; i += 32) {
        const long long k = s0 + i;
        s_d[i] = (k >= 0 && k < len) ? p.demod[base + k] : 0.f;
    }
    __syncwarp();
    if (lane != 0) return;
    int code = -1;
    const double thr = p.thr_rel;
    auto d = [&](long long k) -> double {
        const long long r = k - s0;
        return (double)((r >= 0 && r < VBI_STAGE) ? s_d[r] : p.demod[base + k]);
    };
    // lddutils.calczc with edge='both'; false <=> None (or the IndexError the reference's caller turns into None)
    auto zc = [&](double start_f, int count, double* out) -> bool {
        const long long start = (long long)start_f;
        if (start < 0 || start >= len) return false;
        long long end = start + (long long)count + 1;
        if (end > len) end = len;
        const bool rising = d(start) < thr;
        long long x = -1;
        for (long long k = start; k < end; ++k) {
            const double v = d(k);
            if (rising ? (v >= thr) : (v <= thr)) { x = k; break; }
        }
        if (x <= 0) return false;
        const double a = d(x - 1) - thr, b = d(x) - thr;
        *out = (double)(x - 1) + (-a / (-a + b));
        return true;
    };
    double cur;
    bool have = zc(linestart + 2.0 * fq, (int)(12.0 * fq), &cur);
    int nz = 0;
    unsigned bits = 0;
    double prev = 0.0, gmin = 1e300, gmax = -1e300;
    bool oob = false;
    while (have && nz < 64) {
        const long long bi = (long long)(cur - 0.5 * fq);
        if (bi < 0 || bi >= len) { oob = true; break; }
        if (nz < 24) bits |= (d(bi) < thr ? 1u : 0u) << (23 - nz);
        if (nz > 0) { const double g = (cur - prev) / fq; gmin = g < gmin ? g : gmin; gmax = g > gmax ? g : gmax; }
        prev = cur;
        ++nz;
        const double nxt = cur + 1.9 * fq;
        if ((long long)nxt >= len) { oob = (long long)nxt >= 0; break; }     // data[start_offset] raises IndexError in the reference
        have = zc(nxt, (int)(0.2 * fq), &cur);
    }
    if (!oob && nz == 24 && gmin > 1.85 && gmax < 2.15) code = (int)bits;
    p.codes[(size_t)f * 4 + li] = code;
}

}  // namespace ldd

using namespace ldd;

extern "C" int ldd_vbi_decode(ldd_handle* h, const float* demod_dev, long long n, const long long* base_dev,
                              const long long* winlen_dev, const double* linelocs_dev, int ll_stride, int nfields,
                              const int* lines, int nlines, int* codes_dev, void* stream) {
    if (!h || !demod_dev || !linelocs_dev || !lines || !codes_dev || nlines < 1 || nlines > 4) return LDD_EINVAL;
    if (nfields <= 0) return LDD_OK;
    VbiParams p;
    p.demod = demod_dev; p.n = n; p.base = base_dev; p.winlen = winlen_dev; p.linelocs = linelocs_dev; p.ll_stride = ll_stride;
    for (int i = 0; i < 4; ++i) p.lines[i] = i < nlines ? lines[i] : 0;
    p.nlines = nlines;
    p.freq = h->cfg.freq_hz / 1e6;
    p.thr_rel = h->cfg.hz_ire * 50.0;
    p.codes = codes_dev;
    LDD_LAUNCH(vbi_kernel, dim3(nlines, nfields), dim3(32), 0, (cudaStream_t)stream, p);
    return launch_status(h, "vbi_kernel");
}
