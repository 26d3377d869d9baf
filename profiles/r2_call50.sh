#!/bin/bash
# Round 2 (session 2): the GPU suite on the alternative paths (per-proposal scatter K1'; per-bin pool fold; NMS merge sort).
O=gpurun_out/r2c50
mkdir -p $O
C2D_ROI_TILES=0 timeout 900 python -m pytest tests -m gpu -x -q -k "not tile_owner" > $O/pytest_scatter.log 2>&1; tail -2 $O/pytest_scatter.log
C2D_ROI_FOLD_ROUTED=0 C2D_NMS_MERGE_SORT=1 timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_fold_sort.log 2>&1; tail -2 $O/pytest_fold_sort.log
