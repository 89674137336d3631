#!/bin/bash
# fresh source-level captures of the non-headline conversion directions (columnar->interleaved, C1 both semantics, LAS egress)
mkdir -p gpurun_out
rm -f gpurun_out/prof_directions_r2.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:convert_tiles -s 2 -c 4 -o /tmp/prof_directions_r2b python benchmarks/prof_all.py --points 50000000 --only convert > /dev/null 2>&1
cp /tmp/prof_directions_r2b.ncu-rep gpurun_out/
ls -la gpurun_out/*.ncu-rep
