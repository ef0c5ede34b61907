# builds and runs the isolated phase-A timing probe on the GPU box (nvcc is in the image)
set -e
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -I include -I practical-multi-view_b200/csrc"
nvcc $F -o /tmp/pp tools/phase_a_probe.cu 2>/dev/null; /tmp/pp
