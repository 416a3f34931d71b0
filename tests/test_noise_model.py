"""The certainty test's noise bound (pwknn_host.cpp::noise_bound) against the committed measurement
(profiles/r2_noise_probe.log, written by scripts/noise_probe.py on a B200): the fp32-accumulation bound
c = (D/16 + 4) * 2^-23 must dominate every measured error with headroom, and the bf16-rounding model
sigma = 2.35e-3 / sqrt(D) must match the measured rms on Gaussian rows.  CPU only: keeps the constants in the library,
the document and the evidence file from drifting apart."""

import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _library_constants():
    src = open(os.path.join(ROOT, "pathway_b200", "csrc", "pwknn_host.cpp")).read()
    m = re.search(r"const float c = \(static_cast<float>\(h->dims\) / 16\.f \+ 4\.f\) \* ([0-9.e-]+)f;", src)
    s = re.search(r"static_cast<float>\(h->opt_noise_sigmas\) \* ([0-9.e-]+)f / std::sqrt", src)
    n = re.search(r"int opt_noise_sigmas = (\d+);", src)
    assert m and s and n, "noise_bound() changed shape: update this test and DESIGN.md section 3a together"
    return float(m.group(1)), float(s.group(1)), int(n.group(1))


def test_noise_bound_dominates_the_measured_errors():
    ulp, sigma_coeff, n_sigmas = _library_constants()
    assert abs(ulp - 2.0 ** -23) < 1e-12 and n_sigmas == 8
    rows = [json.loads(ln) for ln in open(os.path.join(ROOT, "profiles", "r2_noise_probe.log")) if ln.startswith("{")]
    assert len(rows) == 12  # D in (64, 768, 1536) x (COS, L2SQ) x (gauss, sparse)
    for r in rows:
        d = r["D"]
        c = (d / 16.0 + 4.0) * ulp
        # default store: only the fp32 accumulation is inexact -- measured max must stay below the bound, with headroom
        assert r["stored_bf16"]["max"] < 0.7 * c, (r["D"], r["metric"], r["data"], r["stored_bf16"]["max"], c)
        if r["data"] == "gauss":
            # exact re-rank store: bf16 rounding noise of raw f32 inputs, Gaussian-like rows
            sigma = sigma_coeff / d ** 0.5
            assert 0.85 < r["raw_f32"]["rms"] / sigma < 1.35, (r["D"], r["metric"], r["raw_f32"]["rms"], sigma)
            # 2.5e7 samples: the largest error seen is ~5.5 sigma; the bound used is n_sigmas = 8
            assert r["raw_f32"]["max"] < n_sigmas * sigma
