#!/bin/bash
# Round 2, last call: the CLI parity cases with long reads (fixture nanopore.bam, synthetic C5, CIGARs of more than 65,535 operations in CG tags) on the shipped build:
# K5's warp-cooperative CIGAR walk through the whole product path.
out=gpurun_out; tag=${1:-r2r}; mkdir -p $out
( timeout 100 python -m pytest tests/test_gpu_cli_parity.py -x -q -k "nanopore or c5" > $out/${tag}_pytest_cli_long_reads.log 2>&1; echo "pytest cli long reads rc=$?"; tail -3 $out/${tag}_pytest_cli_long_reads.log )
