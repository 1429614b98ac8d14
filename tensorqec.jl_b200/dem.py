"""Detector error models: the on-disk input of the circuit-level configs.

Reference: src/decoding/dem.jl:1-6 (`DetectorErrorModel`), :150-160 (`dem2tanner`), :162-164
(`random_error_pattern(dem)`); src/stim_parser/stim_parser.jl:337-375 (`parse_dem_file`, `parse_dem_string`).
Ids are 0-based here: detector D# keeps the number #, observable L# becomes n_detectors + # (the reference stores
# + 1 and largest_detector + # + 1).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .tanner import SimpleTannerGraph


@dataclass
class DetectorErrorModel:
    error_rates: List[float]
    flipped_detectors: List[List[int]]        # per mechanism: detector ids and (offset) observable ids
    detector_list: List[int]
    logical_list: List[int]

    @property
    def n_detectors(self):
        return len(self.detector_list)

    @property
    def n_observables(self):
        return len(self.logical_list)


def parse_dem_string(content: str) -> DetectorErrorModel:
    """stim_parser.jl:342-375.  Only `error(p) ...` lines are read; targets other than D#/L# (e.g. the `^`
    separator of decomposed errors) are ignored, so components are UNIONED (feed undecomposed DEMs);
    `repeat` blocks are rejected."""
    rates, dets, logs = [], [], []
    for line in content.split("\n"):
        s = line.split()
        if not s:
            continue
        if s[0].startswith("error("):
            rates.append(float(s[0][6:-1]))
            d, l = [], []
            for tok in s[1:]:
                if tok.startswith("D"):
                    d.append(int(tok[1:]))
                if tok.startswith("L"):
                    l.append(int(tok[1:]))
            dets.append(d)
            logs.append(l)
        elif s[0].startswith("repeat"):
            raise ValueError("Repeat is not supported, use `circuit.detector_error_model(flatten_loops=True)` "
                             "to flatten the loops in stim")
    if not rates:
        raise ValueError("no error(...) line in the detector error model")
    if any(len(d) == 0 for d in dets):
        # maximum(maximum.(flipped_detectors)) throws on an empty collection in the reference
        raise ValueError("a mechanism without detectors is not supported by the reference parser")
    n_det = max(max(d) for d in dets) + 1
    n_log = max((max(l) + 1 if l else 0) for l in logs)
    flipped = [list(dict.fromkeys(d + [n_det + x for x in l])) for d, l in zip(dets, logs)]
    return DetectorErrorModel(rates, flipped, list(range(n_det)), list(range(n_det, n_det + n_log)))


def parse_dem_file(path: str) -> DetectorErrorModel:
    with open(path, "r") as fh:
        return parse_dem_string(fh.read())


def dem2tanner(dem: DetectorErrorModel) -> SimpleTannerGraph:
    """dem.jl:150-160: bits = mechanisms, checks = detectors 0..max(detector_list)."""
    nqb = len(dem.error_rates)
    logical = set(dem.logical_list)
    q2s = [[d for d in v if d not in logical] for v in dem.flipped_detectors]
    n_s = max(dem.detector_list) + 1
    s2q = [[e for e in range(nqb) if i in q2s[e]] for i in range(n_s)]
    return SimpleTannerGraph(nqb, s2q)
