#!/usr/bin/env python
"""Signatures of 3D tracks dumped from the UNMODIFIED reference (TrackGenerator3D through
oracle/_ref/ref_driver --dump-tracks), for tests/test_trackgen3d.py.

The dumps are 3-150 MB each, so only their signatures are committed: sizes, SHA-256 of every
integer array (links, flags, boundary conditions, segment offsets, FSR ids renumbered in order
of first appearance), and sums / extrema of the floating-point ones.
Needs /root/reference (run in the CPU container):  python tests/golden/make_trackgen3d_signatures.py
"""
import hashlib, json, os, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver")

# name: (model, azim, spacing, polar, z spacing, axial layers, quadrature, formation)
CASES = {
    "c5g7_gl": ("c5g7-2d", 4, 1.0, 2, 10.0, 1, "gl", "otf-stacks"),
    "c5g7_equal_angle_3layers": ("c5g7-2d", 8, 0.8, 4, 6.0, 3, "equal-angle", "otf-tracks"),
    "lattice_equal_weight": ("simple-lattice", 8, 0.3, 6, 0.7, 1, "equal-weight", "otf-stacks"),
    "pin_cell_ty": ("pin-cell", 16, 0.2, 4, 0.3, 1, "ty", "otf-tracks"),
}
INT_KEYS = ["trk_azim", "trk_polar", "trk_xy", "trk_next_fwd", "trk_next_bwd", "trk_flags", "trk_bc_fwd",
            "trk_bc_bwd", "trk_seg_offset", "seg_fsr"]


def signature(ft):
    """Works on a FlatTracks from either side; FSR ids are canonicalised first."""
    from openmoc_b200.synth import _renumber_by_discovery
    a = dict(ft.arrays)
    _renumber_by_discovery(a)
    sig = {"n_tracks": int(ft.n_tracks), "n_segments": int(ft.n_segments), "n_fsrs": int(a["fsr_volume"].size)}
    for k in INT_KEYS:
        sig[k] = hashlib.sha256(np.ascontiguousarray(a[k]).astype("<i8").tobytes()).hexdigest()
    sig["seg_length_sum"] = float(np.sum(a["seg_length"]))
    sig["seg_length_head"] = [float(x) for x in a["seg_length"][:16]]
    sig["fsr_volume_head"] = [float(x) for x in a["fsr_volume"][:16]]
    sig["fsr_volume_sum"] = float(np.sum(a["fsr_volume"]))
    sig["fsr_mat"] = hashlib.sha256(np.ascontiguousarray(a["fsr_mat"]).astype("<i8").tobytes()).hexdigest()
    sig["quad_weight"] = [float(x) for x in a["quad_weight"]]
    sig["quad_sin_theta"] = [float(x) for x in a["quad_sin_theta"]]
    return sig


def main():
    from openmoc_b200.trackfile import read_trackfile
    out = {}
    with tempfile.TemporaryDirectory() as td:
        for name, (model, az, sp, pol, zs, nax, quad, form) in CASES.items():
            trk = os.path.join(td, name + ".b2trk")
            cmd = [DRIVER, "--model", model, "--dims", "3", "--azim", str(az), "--polar", str(pol), "--spacing", str(sp),
                   "--zspacing", str(zs), "--formation", form, "--quad", quad, "--mode", "none", "--quiet",
                   "--dump-tracks", trk]
            if nax > 1:                       # the CMFD mesh is what cuts the reference's FSRs axially
                nxy = {"c5g7-2d": 51, "simple-lattice": 4, "pin-cell": 1}[model]
                cmd += ["--cmfd", f"{nxy}x{nxy}x{nax}"]
            subprocess.check_call(cmd, cwd=td, stdout=subprocess.DEVNULL)
            out[name] = {"case": [model, az, sp, pol, zs, nax, quad, form], "sig": signature(read_trackfile(trk))}
            print(name, out[name]["sig"]["n_tracks"], out[name]["sig"]["n_segments"])
    json.dump(out, open(os.path.join(HERE, "trackgen3d_signatures.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
