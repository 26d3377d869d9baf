#!/bin/bash
# Round 2 (session 2): HBM bandwidth by direction (write-only / read-only / copy) on the box.
timeout 300 python profiles/microbench_hbm_rw.py | tail -1
