#!/bin/bash
# scripts/build_variant.sh <name> [-DFLAG=..]...  ->  build/lib<name>.so  (A/B builds; select with PARAMOL_B200_LIB)
name=$1; shift
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" \
  paramol_b200/csrc/pm_kernels.cu paramol_b200/csrc/pm_fit_kernels.cu paramol_b200/csrc/pm_grad_kernels.cu paramol_b200/csrc/pm_delta_kernels.cu paramol_b200/csrc/pm_spec.cu paramol_b200/csrc/pm_comm.cu paramol_b200/csrc/pm_lls_kernels.cu paramol_b200/csrc/pm_egemm_kernels.cu -o build/lib$name.so
