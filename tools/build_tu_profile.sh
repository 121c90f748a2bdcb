#!/bin/bash
# rebuilds k_train_umma.o with the clock64 timeline enabled and relinks libtbnn.so (undo: touch the .cu and rebuild)
cd "$(dirname "$0")/../tensorbnn_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DTBNN_TU_PROFILE -c k_train_umma.cu -o k_train_umma.o || exit 1
nvcc -shared -o ../libtbnn.so api.o k_main.o k_wide.o k_wide2.o k_sweep_umma.o k_train_umma.o k_hyper.o k_predict.o k_predict_umma.o k_adapter.o -ldl
