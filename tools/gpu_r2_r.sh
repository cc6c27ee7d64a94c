#!/bin/bash
# round 2, call R: parity table + list of divergent trajectories with the final kernels
mkdir -p gpurun_out
python tools/parity_report.py --oracle-n 20000 --divergent gpurun_out/r02_divergent.json > gpurun_out/r02_parity_table.txt 2> gpurun_out/r02_parity.err; echo "parity rc=$?"
tail -25 gpurun_out/r02_parity_table.txt
