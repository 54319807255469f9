#!/bin/bash
# Build the standalone decoder-stream harness (tests/cuda/dec_stream_test.cu) for sm_100a.
set -e
cd "$(dirname "$0")"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fopenmp \
  -o dec_stream_test dec_stream_test.cu ../../orca_b200/csrc/conv2d_stream.cu ../../orca_b200/csrc/dec_glue.cu -lgomp
# event-trace variant (CTA 20): ./dec_stream_trace trace2 > trace.txt
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fopenmp -DDS_TRACE=20 \
  -o dec_stream_trace dec_stream_test.cu ../../orca_b200/csrc/conv2d_stream.cu ../../orca_b200/csrc/dec_glue.cu -lgomp
