#!/bin/bash
# Round 2: chain groups (gridDim.z) of the TMA stencil, loss-only variant.  usage: r2_groups_sweep.sh C N reps G...
cd "$(dirname "$0")/.."
C=$1; N=$2; R=$3; shift 3
echo "model's own choice:"; python profiles/stencil_only.py $C $N $R 2>&1 | sed -n '1,2p'
for g in "$@"; do echo "G=$g"; GMC_RS_GROUPS=$g python profiles/stencil_only.py $C $N $R 2>&1 | sed -n '1,2p'; done
