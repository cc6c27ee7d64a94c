#!/bin/bash
# Build an experimental variant of the library beside the product one and print the static counts of its C3 / C4 hot loops.
#   tools/variant.sh <tag> [nvcc flags, e.g. -DNI_EXP_FOO=1]   ->  gpurun_exp/libni_<tag>.so (A/B: tools/exp_c3.sh)
tag=$1; shift
obj=/tmp/ni_variant_$tag
mkdir -p $obj gpurun_exp
NI_B200_OBJ_DIR=$obj NI_B200_LIB_OUT=$PWD/gpurun_exp/libni_$tag.so NI_EXTRA_NVCC_FLAGS="$*" python -m numba_integrators_b200.build > /dev/null || exit 1
python tools/sass_count.py $obj/ni_builtin_lorenz63.o _Z8ni_k_runI13NiRhsLorenz63Li1ELi1ELb0ELb0EEv6NiArgs | head -2
python tools/sass_count.py $obj/ni_builtin_vanderpol.o _Z8ni_k_runI14NiRhsVanDerPolLi0ELi0ELb0ELb0EEv6NiArgs | head -2
