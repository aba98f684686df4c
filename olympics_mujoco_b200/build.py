"""Build libom_b200.so in-tree: run the model code generator, then nvcc for sm_100a.

    python -m olympics_mujoco_b200.build [--force]

The shared library lands at ``olympics_mujoco_b200/libom_b200.so`` (git-ignored, travels with gpurun).
nvcc cross-compiles without a GPU, so this also runs in the authoring container.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

from . import codegen, mjcf

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
GEN = CSRC / "gen"
LIB = PKG / "libom_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]

H1_SPEC_JOINTS = ["pelvis_tx", "pelvis_tz", "pelvis_ty", "pelvis_tilt", "pelvis_list", "pelvis_rotation", "back_bkz",
                  "hip_flexion_r", "hip_adduction_r", "hip_rotation_r", "knee_angle_r", "ankle_angle_r",
                  "hip_flexion_l", "hip_adduction_l", "hip_rotation_l", "knee_angle_l", "ankle_angle_l"]


def _write_if_changed(path: Path, text: str):
    if path.exists() and path.read_text() == text:
        return False
    path.write_text(text)
    return True


def generate():
    """(Re)generate csrc/gen/* from the shipped model tables."""
    GEN.mkdir(exist_ok=True)
    h1 = mjcf.load_builtin("unitree_h1")
    a3 = mjcf.load_builtin("stick_figure_a3")
    _write_if_changed(GEN / "fk_unitree_h1.cuh", "#pragma once\n" + codegen.generate_fk(h1, "om_fk_unitree_h1"))
    _write_if_changed(GEN / "fk_stick_figure_a3.cuh", "#pragma once\n" + codegen.generate_fk(a3, "om_fk_stick_figure_a3"))
    _write_if_changed(GEN / "fk_pos_stick_figure_a3.cuh",
                      "#pragma once\n" + codegen.generate_fk_pos(a3, "om_fk_pos_stick_figure_a3"))
    # float64 twin of the position FK: the exact slow path of the A3 threshold decisions (done, target_reached)
    _write_if_changed(GEN / "fk_pos_f64_stick_figure_a3.cuh",
                      "#pragma once\n" + codegen.generate_fk_pos(a3, "om_fk_pos_f64_stick_figure_a3", scalar="double"))
    parts = codegen.split_parts(h1, 3)
    _write_if_changed(GEN / "fk_unitree_h1_parts.cuh", "#pragma once\n" + "".join(
        codegen.generate_fk(h1, f"om_fk_unitree_h1_part{k}", part=p) for k, p in enumerate(parts)))
    _write_if_changed(GEN / "tables_unitree_h1.h", "#pragma once\n" + codegen.generate_tables(h1, "om_tab_h1"))
    _write_if_changed(GEN / "tables_stick_figure_a3.h", "#pragma once\n" + codegen.generate_tables(a3, "om_tab_a3"))
    perm = [int(h1.jnt_qposadr[h1.jnt_names.index(j)]) for j in H1_SPEC_JOINTS]
    txt = ("#pragma once\n// GENERATED: qpos/qvel address of UnitreeH1 observation-spec entry k "
           "(UnitreeH1.py:303-355 minus the arm joints)\n"
           f"__device__ constexpr int OM_H1_PERM[17] = {{{', '.join(map(str, perm))}}};\n"
           f"static const int OM_H1_PERM_HOST[17] = {{{', '.join(map(str, perm))}}};\n")
    _write_if_changed(GEN / "h1_perm.h", txt)
    ids = dict(ROOT=a3.body_id("torso"), HEAD=a3.body_id("head"), LFOOT=a3.body_id("left_foot"),
               RFOOT=a3.body_id("right_foot"), LSITE=a3.site_id("lf_force"), RSITE=a3.site_id("rf_force"))
    txt = ("#pragma once\n// GENERATED: body / site ids the WalkingTask reads (StickFigureA3.py:95-103, "
           "walking_task.py:254-263) and the model's total mass (mj_getTotalmass)\n" +
           "".join(f"constexpr int OM_A3_{k} = {v};\n" for k, v in ids.items()) +
           f"constexpr float OM_A3_TOTAL_MASS = {codegen._f(a3.total_mass)};\n")
    _write_if_changed(GEN / "a3_ids.h", txt)


def sources():
    return sorted(CSRC.glob("*.cu"))


def _digest():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(GEN.glob("*")) +
                    [PKG.parent / "include" / "om_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force=False, verbose=True):
    generate()
    stamp = PKG / ".libom_b200.stamp"
    dig = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == dig:
        return LIB
    objs = []
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [NVCC, *ARCH, *FLAGS, "-I", str(PKG.parent / "include"), "-c", str(src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src.name}\n{out}")
        if p.returncode != 0:
            failed = True
    (objdir / "ptxas.log").write_text("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed (see output above)")
    cmd = [NVCC, *ARCH, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart"]
    subprocess.check_call(cmd)
    stamp.write_text(dig)
    if verbose:
        print(f"built {LIB} ({LIB.stat().st_size / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
