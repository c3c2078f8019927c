#!/bin/bash
# Stages the files tests/harness/run_reference_scripts.py needs from the read-only reference mount into baseline/_ref/
# (git-ignored, NOT gpurun-ignored: it travels to the GPU box, it never enters the history).
set -e
SRC=${1:-/root/reference}
DST="$(dirname "$0")/../baseline/_ref"
mkdir -p "$DST/duffing" "$DST/8x8_cloth_swing_xyz"
cp "$SRC"/regressors.py "$SRC"/dynamical_systems.py "$SRC"/benchmark_lqr_hjb.py "$SRC"/benchmark_lqr_classic.py "$SRC"/benchmark_lqr_cloth.py "$DST"/
cp "$SRC"/duffing/duffing_{x,y,u}_forced.csv "$SRC"/duffing/duffing_{x,y}_unforced.csv "$SRC"/duffing/all_rmses_nystrom_double_dataset.csv "$DST"/duffing/
cp "$SRC"/8x8_cloth_swing_xyz/state_samples_cloth_swing_*.csv "$SRC"/8x8_cloth_swing_xyz/input_samples_cloth_swing_*.csv "$DST"/8x8_cloth_swing_xyz/
du -sh "$DST"
