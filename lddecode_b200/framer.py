"""Framer and downscale_audio: host-side mirrors of the reference's frame assembly
(lddecode_core.py:431-484, 1193-1378) on top of the device-backed RFDecode / Field classes.

These are the callers either side of the hot path: control logic
that is O(1) per field (field pairing, VBI merge, MTF adaptation, seek) plus the 48 kHz audio
resample along the line positions.  They keep the reference's names, arguments and return values so
that lddecode.py's main loop runs unchanged on them.
"""
import copy
import io

import numpy as np

from . import field as F


def downscale_audio(audio, lineinfo, rf, linecount, timeoffset=0, freq=48000.0, scale=64):
    """lddecode_core.py:431-484.  `audio` is the phase-2 record array of the field's window.

    Returns (int16 interleaved L/R samples, time offset to carry into the next field).  Like the
    reference it assumes a total decimation of `scale` (64: the 40 MSPS default)."""
    frametime = (rf.SysParams['line_period'] * linecount) / 1000000
    soundgap = 1 / freq
    arange = np.arange(timeoffset, frametime + soundgap, soundgap, dtype=np.double)
    lineinfo = np.asarray(lineinfo, dtype=np.float64)
    linenum = ((arange * 1000000) / rf.SysParams['line_period']) + 1
    li = linenum.astype(np.int64)                     # np.int(linenum): truncation
    cur = lineinfo[li]
    nxt = np.where(li + 1 < len(lineinfo), lineinfo[np.minimum(li + 1, len(lineinfo) - 1)], cur + rf.linelen)
    sampleloc = cur + (nxt - cur) * (linenum - np.floor(linenum))
    swow = (nxt - cur) / rf.linelen
    locs = sampleloc / scale
    n = len(arange) - 1
    idx = locs[:n].astype(np.int64)
    left = np.asarray(audio['audio_left'], dtype=np.float64)[idx] * swow[:n] - rf.SysParams['audio_lfreq']
    right = np.asarray(audio['audio_right'], dtype=np.float64)[idx] * swow[:n] - rf.SysParams['audio_rfreq']
    out = np.zeros(2 * n, dtype=np.int32)
    out[0::2] = np.round(left * 32767 / 150000).astype(np.int32)
    out[1::2] = np.round(right * 32767 / 150000).astype(np.int32)
    out16 = np.zeros(2 * n, dtype=np.int16)
    np.clip(out, -32766, 32766, out=out16)
    return out16, arange[-1] - frametime


def field_audio(f):
    """What Field.downscale(audio=True) adds (lddecode_core.py:809-810): sets f.dsaudio / f.audio_next_offset."""
    rf = f.rf
    if not rf.decode_analog_audio or f.audio_rec is None:
        return None
    f.dsaudio, f.audio_next_offset = downscale_audio(f.audio_rec, f.linelocs, rf, f.linecount, f.audio_next_offset)
    return f.dsaudio


class Framer:
    """lddecode_core.py:1193-1334 with FieldClass = the device-backed FieldNTSC / FieldPAL."""

    def __init__(self, rf, full_decode=True):
        self.rf = rf
        self.full_decode = full_decode
        if rf.system == 'PAL':
            self.FieldClass, self.readlen, self.outlines, self.clvfps = F.FieldPAL, 1000000, 625, 25
        else:
            self.FieldClass, self.readlen, self.outlines, self.clvfps = F.FieldNTSC, 1000000, 525, 30
        if not full_decode:
            self.FieldClass = F.Field
        self.outwidth = rf.SysParams['outlinelen']
        self.audio_offset = 0
        self.mtf_level = 1
        self.vbi = None

    def readfield(self, infile, sample, fieldcount=0):
        readsample = sample
        while True:
            raw = self.rf.demod_raw(infile, readsample, self.readlen, self.mtf_level)
            if raw is None:
                return None, None, None
            f = self.FieldClass(self.rf, raw, 0, audio_offset=self.audio_offset)
            nextsample = readsample + f.nextfieldoffset
            if not f.valid:
                if len(f.peaklist) < 100:
                    nextsample = readsample + (self.rf.freq_hz * 10)
                elif len(f.vsyncs) == 0:
                    nextsample = readsample + (self.rf.freq_hz * 1)
                readsample = nextsample
            else:
                if self.full_decode and self.rf.decode_analog_audio:
                    f.audio_rec = raw.audio_recarray()
                    field_audio(f)
                return f, readsample, nextsample

    def mergevbi(self, fields):
        merged = copy.copy(fields[0].vbi)
        for k in merged.keys():
            if fields[1].vbi[k] is not None:
                merged[k] = fields[1].vbi[k]
        if merged['seconds'] is not None:
            merged['framenr'] = merged['minutes'] * 60 * self.clvfps + merged['seconds'] * self.clvfps + merged['clvframe']
        return merged

    def formatoutput(self, fields):
        W = self.outwidth
        linecount = min(fields[0].linecount, fields[1].linecount) * 2
        combined = np.zeros((W * self.outlines), dtype=np.uint16)
        ol = fields[0].outlinelen
        for i in range(0, linecount, 2):
            cur = i // 2
            combined[i * W:(i + 1) * W] = fields[0].dspicture[cur * ol:cur * ol + W]
            combined[(i + 1) * W:(i + 2) * W] = fields[1].dspicture[cur * ol:cur * ol + W]
        lf = int(np.argmax([fields[0].linecount, fields[1].linecount]))
        cur = linecount // 2
        combined[linecount * W:(linecount + 1) * W] = fields[lf].dspicture[cur * ol:cur * ol + W]
        return combined

    def readframe(self, infile, sample, firstframe=False, CAV=False):
        fieldcount = 0
        fields = [None, None]
        audio = []
        f = None
        while fieldcount < 2:
            f, readsample, nextsample = self.readfield(infile, sample, fieldcount)
            if f is not None:
                if f.istop:
                    fields[0] = f
                else:
                    fields[1] = f
                if ((not CAV and (f.istop == self.rf.SysParams['topfirst'])) or
                        (CAV and (f.vbi['framenr'] or f.vbi['minutes']))):
                    fieldcount = 1
                elif fieldcount == 1:
                    fieldcount = 2
                if (fieldcount or not firstframe) and f.dsaudio is not None:
                    audio.append(f.dsaudio)
            elif readsample is None:
                return None, None, None, None
            sample = nextsample
        if len(audio):
            conaudio = np.concatenate(audio)
            self.audio_offset = f.audio_next_offset
        else:
            conaudio = None
        combined = self.formatoutput(fields) if self.full_decode else None
        self.vbi = self.mergevbi(fields)
        if not f.vbi['isclv'] and f.vbi['framenr'] is not None:
            newmtf = max(1 - (f.vbi['framenr'] / 10000), 0)
            oldmtf = self.mtf_level
            self.mtf_level = newmtf
            if np.abs(newmtf - oldmtf) > .1:
                return self.readframe(infile, sample, firstframe, CAV)
        return combined, conaudio, sample, fields


def findframe(infile, rf, target, nextsample=0):
    """lddecode_core.py:1338-1378: locate the sample number of the target frame."""
    framer = Framer(rf, full_decode=False)
    samples_per_frame = int(rf.freq_hz / rf.SysParams['FPS'])
    framer.vbi = {'framenr': None}
    iscav = False
    retry = 5
    rv = None
    tolerance = 0
    while framer.vbi['framenr'] is None and retry:
        rv = framer.readframe(infile, nextsample, CAV=False)
        if framer.vbi['isclv']:
            tolerance = 1
        else:
            tolerance = 0
            iscav = True
        nextsample = rv[2] + (rf.freq_hz * 10)
        retry -= 1
    if retry == 0 and framer.vbi['framenr'] is None:
        print("SEEK ERROR: Unable to find a usable frame")
        return None
    retry = 5
    while np.abs(target - framer.vbi['framenr']) > tolerance and retry:
        offset = samples_per_frame * (target - 1 - framer.vbi['framenr'])
        nextsample = rv[2] + offset
        rv = framer.readframe(infile, nextsample, CAV=iscav)
        retry -= 1
    if np.abs(target - framer.vbi['framenr']) > tolerance:
        print("SEEK WARNING: seeked to frame {0} instead of {1}".format(framer.vbi['framenr'], target))
    return nextsample
