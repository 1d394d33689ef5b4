#!/usr/bin/env python
"""Generates tests/golden/golden_small.npz and reference_known_answers.json.

The reference (Rust) cannot be compiled or imported here, so the vectors are produced by the
CPU oracle and ACCEPTED ONLY IF they agree with numpy's pocketfft under the sign/packing
mapping of SURVEY.md section 8a/8c (asserted below).  The JSON file restates the known answers
held by the reference's own unit tests (file:line cited per entry; only the rows SURVEY.md
section 4 marks P = satisfiable).
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def rel(a, b):
    return np.linalg.norm(np.ravel(a) - np.ravel(b)) / np.linalg.norm(np.ravel(b))


def c(z):
    return z[0::2] + 1j * z[1::2]


def main():
    out = {}
    for nn in (8, 64, 1024):
        x = O.fill_uniform(1001, 0, 2 * nn)
        for s in (1, -1):
            y = O.four1(x.copy(), nn, s)
            ref = np.fft.ifft(c(x)) * nn if s == 1 else np.fft.fft(c(x))
            assert rel(c(y), ref) < 2e-15
            out[f"four1_{nn}_{'p' if s == 1 else 'm'}"] = y
        out[f"four1_{nn}_in"] = x
    for shape in ((4, 8, 2), (8, 16), (16, 4, 8)):
        n = int(np.prod(shape))
        x = O.fill_uniform(1003, 0, 2 * n)
        tag = "x".join(map(str, shape))
        for s in (1, -1):
            y = O.fourn(x.copy(), list(shape), s)
            ref = np.fft.ifftn(c(x).reshape(shape)) * n if s == 1 else np.fft.fftn(c(x).reshape(shape))
            assert rel(c(y), ref.ravel()) < 2e-15
            out[f"fourn_{tag}_{'p' if s == 1 else 'm'}"] = y
        out[f"fourn_{tag}_in"] = x
    for n in (8, 256, 2048):
        x = O.fill_uniform(1004, 0, n)
        y = O.realft(x.copy(), n, 1)
        F = np.conj(np.fft.rfft(x))
        pk = np.empty(n)
        pk[0], pk[1] = F[0].real, F[n // 2].real
        pk[2::2], pk[3::2] = F[1:n // 2].real, F[1:n // 2].imag
        assert rel(y, pk) < 2e-15
        z = O.realft(y.copy(), n, -1)
        assert rel(z * 2 / n, x) < 4e-15
        out[f"realft_{n}_in"], out[f"realft_{n}_fwd"], out[f"realft_{n}_rt"] = x, y, z
    for shp in ((8, 8, 8), (4, 16, 8), (2, 4, 32)):
        x = O.fill_uniform(1006, 0, int(np.prod(shp))).reshape(shp)
        d, s = O.rlft3(x.copy(), np.zeros((shp[0], 2 * shp[1])), 1)
        F = np.conj(np.fft.rfftn(x))
        dd = d.reshape(shp[0], shp[1], shp[2] // 2, 2)
        ss = s.reshape(shp[0], shp[1], 2)
        assert rel(dd[..., 0] + 1j * dd[..., 1], F[..., :shp[2] // 2]) < 2e-15
        assert rel(ss[..., 0] + 1j * ss[..., 1], F[..., shp[2] // 2]) < 2e-15
        tag = "x".join(map(str, shp))
        out[f"rlft3_{tag}_in"], out[f"rlft3_{tag}_data"], out[f"rlft3_{tag}_speq"] = x, d, s
    a = O.fill_uniform(1004, 0, 128)
    r = O.fill_uniform(1005, 0, 9) / 64.0
    rc, y = O.convlv(a, r, 1, 0)
    p = O.pad_response(r, 128, 0)
    assert rc == 0 and rel(y, np.fft.irfft(np.fft.rfft(a) * np.fft.rfft(p), 128)) < 4e-15
    out["convlv_128_9_in"], out["convlv_128_9_resp"], out["convlv_128_9_out"] = a, r, y
    b = O.fill_uniform(1011, 0, 128)
    rc, y = O.correl(a, b)
    assert rc == 0 and rel(y, np.fft.irfft(np.fft.rfft(a) * np.conj(np.fft.rfft(b)), 128)) < 4e-15
    out["correl_128_a"], out["correl_128_b"], out["correl_128_out"] = a, b, y
    np.savez_compressed(os.path.join(HERE, "golden_small.npz"), **out)

    known = {
        "four1_round_trip": {"cite": "FFT_1.rs:246-267", "n": 1024, "signal": "sin(2*pi*5*t)+0.5*cos(2*pi*20*t), t=i/1024",
                             "scale": "1/N", "abs_tol": 1e-10},
        "convlv_basic": {"cite": "Convolve.rs:347-360,445-454", "data": [1, 2, 3, 4], "respns": [1, 1], "isign": 1,
                         "expect_at": {"1": 3.0, "2": 5.0, "3": 7.0}, "abs_tol": 1e-10,
                         "note": "index 0 expects 1.0 in the reference test, which circular convolution cannot give (5.0)"},
        "convlv_errors": {"cite": "Convolve.rs:406-424",
                          "cases": [{"data": [], "respns": [1.0], "isign": 1, "err": "EmptyInput"},
                                    {"data": [1.0, 2.0], "respns": [1.0, 2.0, 3.0], "isign": 1, "err": "ResponseTooLong"},
                                    {"data": [1.0, 2.0], "respns": [1.0], "isign": 0, "err": "InvalidIsign"}]},
        "correl_basic": {"cite": "Correlation.rs:407-418", "a": [1, 2, 3, 4], "b": [1, 2, 3, 4], "expect_at": {"0": 30.0}},
        "correl_batch": {"cite": "Correlation.rs:465-478", "pairs": [[[1, 2], [1, 2]], [[3, 4], [3, 4]]], "expect0": [5.0, 25.0]},
        "autocorrel": {"cite": "Correlation.rs:481-491", "a": [1, 2, 1, 2], "expect_at": {"0": 10.0}},
        "correl_direct_small": {"cite": "Correlation.rs:494-502", "a": [1, 2], "b": [1, 2], "expect": [5.0, 2.0]},
        "correl_errors": {"cite": "Correlation.rs:453-462",
                          "cases": [{"a": [], "b": [1.0], "err": "EmptyInput"}, {"a": [1.0, 2.0], "b": [1.0], "err": "LengthMismatch"}]},
        "fourn_validation": {"cite": "Fourn.rs:367-378,467-476",
                             "cases": [{"nn": [8, 8], "ndim": 2, "isign": 1, "ok": True}, {"nn": [1], "ndim": 1, "isign": 1, "ok": False},
                                       {"nn": [8], "ndim": 1, "isign": 0, "ok": False}]},
    }
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(known, f, indent=1)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
