"""Bisects a chain scene (events + a > 2 s input gap) effect by effect: GPU vs oracle (debug aid)."""
import os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import phonic_b200
from conftest import ORACLE_LIB
from phonic_b200._capi import CApi
from phonic_b200.player import Player, ChorusEffect, CompressorEffect, Eq5Effect, FilePlaybackOptions, FilterEffect
from phonic_b200 import workloads as W
from scenes import tone
SRX = 48000
apis = {"gpu": phonic_b200.load_api(), "oracle": CApi(ORACLE_LIB, "po_")}
def scene(api, which, events):
    p = Player(api, SRX)
    b = p.upload_buffer(tone(20000, 44100, seed=41), 44100)
    fx = {}
    if "f" in which: fx["f"] = p.add_effect(FilterEffect(0, 3000.0, 0.707))
    if "e" in which: fx["e"] = p.add_effect(Eq5Effect())
    if "c" in which: fx["c"] = p.add_effect(CompressorEffect())
    if "r" in which: fx["r"] = p.add_effect(ChorusEffect())
    if events:
        if "e" in fx: fx["e"].set_parameter("gan2", 5.0, 0); fx["e"].set_parameter("gan4", -4.0, 15555)
        if "f" in fx: fx["f"].set_parameter("cuto", 900.0, 7001); fx["f"].set_parameter("cuto", 4000.0, int(3.1 * SRX) + 13)
        if "c" in fx: fx["c"].set_parameter("thrs", -24.0, 23456)
        if "r" in fx: fx["r"].set_parameter("rate", 1.5, int(3.4 * SRX) + 5)
    p.play_file_source(b, FilePlaybackOptions(volume=0.6)); p.play_file_source(b, FilePlaybackOptions(volume=0.6), start_time=int(3.0 * SRX) + 321)
    out = p.render(W.frames_for(4, SRX)); p.close()
    return out
for which in ("f", "e", "c", "r", "fe", "fec", "fecr"):
    for events in (False, True):
        a, c = scene(apis["gpu"], which, events), scene(apis["oracle"], which, events)
        d = np.abs(a - c).max(axis=1)
        big = np.flatnonzero(d > 1e-5)
        print(which, "events" if events else "plain ", "max %.2e" % d.max(), "first >1e-5:", int(big[0]) if big.size else None)
