"""phonic_b200 -- B200-native offline renderer for emuell/phonic's mixer graph.

The product is the CUDA shared library `phonic_b200/csrc/libphonic_b200.so` (C-ABI in
include/phonic_b200.h); this package is the Python host-side mirror of the reference's
Player/handle API on top of it. There is NO CPU fallback: loading fails loudly when the CUDA
library has not been built, and rendering fails loudly without a CUDA device.
"""
from __future__ import annotations

import os

from . import _capi
from ._capi import CApi
from .player import (AhdsrParameters, ChorusEffect, CompressorEffect, DelayEffect, EffectHandle, Eq5Effect,
                     FilePlaybackHandle, FilePlaybackOptions, FilterEffect, GeneratorPlaybackHandle,
                     GeneratorPlaybackOptions, MixerHandle, PhonicError, Player, ReverbEffect)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB200_LIB") or os.path.join(_HERE, "csrc", "libphonic_b200.so")  # PB200_LIB: A/B builds of the same library

_api = None


def load_api() -> CApi:
    """Load the CUDA product library. Raises if it is not built (no fallback of any kind)."""
    global _api
    if _api is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(nvcc, sm_100a). phonic_b200 has no CPU fallback.")
        _api = CApi(LIB_PATH, "pb200_")
    return _api


def new_player(sample_rate: int = 48000, device_ordinal: int = -1, **kw) -> Player:
    """Player::new(WavOutput::open_with_specs(.., sample_rate, 2, ..)) on the CUDA renderer."""
    return Player(load_api(), sample_rate=sample_rate, device_ordinal=device_ordinal, **kw)
