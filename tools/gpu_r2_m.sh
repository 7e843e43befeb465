#!/bin/bash
# Round-2 visit M (1 GPU): per-shape eager tables of the step with and without the LayerNorm fold (which GEMMs pay for it)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 200 python tools/step_profile.py --reps 7 > gpurun_out/step_profile_lnfold.txt 2>&1; echo "fold rc=$?"
MVD_NO_LN_FOLD=1 timeout 200 python tools/step_profile.py --reps 7 > gpurun_out/step_profile_lnpass.txt 2>&1; echo "pass rc=$?"
head -2 gpurun_out/step_profile_lnfold.txt gpurun_out/step_profile_lnpass.txt
