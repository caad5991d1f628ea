"""Development aid: run the NTSC field chain on the GPU with a sync after every launch."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from lddecode_b200 import _lib, field as F, rfdecode

g = dict(np.load(os.path.join(ROOT, "tests/golden/ntsc.npz")))
rf = rfdecode.RFDecode(float(g["fs_mhz"]), "NTSC", int(g["blocklen"]))
be = rf._be
cap = g["capture"]
dd = rf.demod_device(be.to_device(cap), _lib.FMT_U8, 0, len(cap), 0, int(g["demod_length"]), 1)
torch.cuda.synchronize(); print("demod ok")
pk, vl = F.sync_peaks_device(rf, dd.planes["demod_sync"], dd.length, 0)
print("peaks", len(pk), np.array_equal(pk, g["field_peaklist"]))
info, ll1, bad = F.locate(rf, pk, vl, dd.length, 0)
print("stage", info.stage, info.linecount)
batch = F.FieldBatch(rf, 1)
batch.linecount[0] = info.linecount; batch.winlen[0] = dd.length; batch.linelocs1[0] = ll1; batch.linebad[0] = bad
lib = be.lib
n = 1; LL = F.LL_STRIDE
d_base = be.to_device(batch.base); d_win = be.to_device(batch.winlen); d_lc = be.to_device(batch.linecount)
d_l1 = be.to_device(batch.linelocs1.reshape(-1)); d_bad = be.to_device(batch.linebad.reshape(-1))
d_l2 = be.empty(LL, np.float64); d_bad2 = be.empty(LL, np.uint8); d_status = be.zeros(1, np.int32)
st = be.stream()
def chk(name, rc):
    print(name, "rc", rc, (lib.ldd_last_error(rf._h) or b"").decode() if rc else "")
    try:
        torch.cuda.synchronize(); print("   sync ok")
    except Exception as e:
        print("   SYNC ERROR", e); sys.exit(1)
chk("hsync", lib.ldd_refine_hsync(rf._h, be.ptr(dd.planes['demod_05']), dd.length, be.ptr(d_base), be.ptr(d_win), be.ptr(d_lc), n, LL, be.ptr(d_l1), be.ptr(d_bad), be.ptr(d_l2), be.ptr(d_bad2), be.ptr(d_status), st))
d_l3 = be.empty(LL, np.float64); d_l4 = be.empty(LL, np.float64); d_bl = be.empty(LL, np.float32)
chk("burst1", lib.ldd_refine_burst(rf._h, be.ptr(dd.planes['demod_burst']), dd.length, be.ptr(d_base), be.ptr(d_lc), n, LL, be.ptr(d_l2), be.ptr(d_l3), be.ptr(d_bl), be.ptr(d_status), st))
chk("burst2", lib.ldd_refine_burst(rf._h, be.ptr(dd.planes['demod_burst']), dd.length, be.ptr(d_base), be.ptr(d_lc), n, LL, be.ptr(d_l3), be.ptr(d_l4), be.ptr(d_bl), be.ptr(d_status), st))
W = rf.SysParams['outlinelen']; stride = 263 * W
d_pic = be.empty(stride, np.uint16)
chk("tbc", lib.ldd_tbc_fields(rf._h, be.ptr(dd.planes['demod']), dd.length, float(rf.SysParams['ire0']), be.ptr(d_base), be.ptr(d_l4), LL, be.ptr(d_lc), n, int(info.linecount), 1, -12.8, W, 1, 1, be.ptr(d_pic), stride, be.ptr(d_bl), 1.45, be.ptr(d_status), st))
print("status", be.to_host(d_status))
d = be.to_host(d_pic)[:info.linecount * W].astype(int) - g["field_dspicture"].astype(int)
print("pic maxdiff", np.abs(d).max(), np.count_nonzero(d))
