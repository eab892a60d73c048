#!/bin/bash
# runtime knobs of the queue wavefront (no rebuild): pools x trace blocks per SM x slot cap
spp=${1:-128}; k=${2:-16}
for pools in 1 2 3 4; do for tb in 2 3 4; do for slots in 1048576 2097152 4194304; do
  echo -n "pools=$pools trace_blocks=$tb slots=$slots: "
  VIMG_QW_POOLS=$pools VIMG_QW_TRACE_BLOCKS=$tb VIMG_QW_SLOTS=$slots python scripts/streams_sweep.py disney_spheres $spp $k 2>&1 | tail -1 | cut -d: -f2 | cut -d, -f1,3
done; done; done
