#!/bin/bash
# torchrun wrapper: rank 0 runs under ncu (full set, the structured stage kernel only), the other ranks plainly.
# Use with T8B200_SYNC=peer: a replayed stage kernel only reads peer memory that cannot change while the peers wait in
# the next barrier, and a replayed barrier kernel re-stores the same epoch.
# usage: torchrun ... --no-python tools/ncu_rank0.sh OUT.ncu-rep python bench.py ARGS
out=$1; shift
if [ "${LOCAL_RANK:-0}" = "0" ]; then
  exec ncu --set full --section Nvlink --section Nvlink_Tables --clock-control none --import-source on \
       -k regex:structured_stage_kernel --launch-skip ${NCU_SKIP:-12} -c ${NCU_COUNT:-3} -o "$out" "$@"
else
  exec "$@"
fi
