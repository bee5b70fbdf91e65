#!/bin/bash
P=scripts/_bin/umma_probe
run() { timeout 60 $P "$@" 2>&1 || echo "FAILED: $@"; }
# control: f16 K-major, f16 MN-major SW128
run A K 2 0 16 1024 0 128 1
run A MN 2 0 4096 1024 0 128 1
run A MN 2 0 1024 4096 0 128 1
run B MN 2 0 4096 1024 0 64 1
# tf32 MN-major with the other layout types
run A MN 1 0 4096 1024 0 128
run A MN 1 0 1024 4096 0 128
run A MN 4 0 4096 1024 0 128
run A MN 6 0 4096 1024 0 128
run A MN 3 0 4096 1024 0 128
run A MN 5 0 4096 1024 0 128
run B MN 1 0 4096 1024 0 64
run A MN 1 128 4096 1024 0 128
run A MN 1 384 4096 1024 0 128
