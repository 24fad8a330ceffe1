#!/bin/bash
# Development builds of the shell engine's compile-time variants (all OFF in the shipped libtuna_b200.so), for A/B timing with
# tools/direct_timing.py on a GPU box.  Each is checked for parity on the CPU by tests/test_host_emul.py (same macros, HostPolicy).
#   TUNA_SHELL_WIDE_TERMS   phase 5: 8-byte terms with pre-scaled byte offsets, ping-pong quads (no decode, no register copies)
#   TUNA_SHELL_ASM_UNROLL   phase 4: y operands in registers, m' loop unrolled per trip count
#   TUNA_SHELL_REG_TIERS    128-register instantiation of k_shell_jk_one for class jobs whose shared memory limits occupancy anyway
#   TUNA_MO_DMMA            AO->MO steps on the FP64 tensor cores (mma.sync m8n8k4 f64); check with
#                           TUNA_B200_LIB=build/lib_dmma.so python -m pytest tests/test_mo_transform.py -m gpu && TUNA_B200_LIB=build/lib_dmma.so python tools/mo_quick.py
set -e
cd "$(dirname "$0")/.."
mkdir -p build
build() { TUNA_B200_LIB=$PWD/build/$1 TUNA_B200_NVCC_EXTRA="$2" python -m tuna_b200.build --force; }
build lib_wide.so  "-DTUNA_SHELL_WIDE_TERMS"
build lib_asm.so   "-DTUNA_SHELL_ASM_UNROLL"
build lib_tiers.so "-DTUNA_SHELL_REG_TIERS"
build lib_v3.so    "-DTUNA_SHELL_WIDE_TERMS -DTUNA_SHELL_ASM_UNROLL -DTUNA_SHELL_REG_TIERS"
build lib_dmma.so  "-DTUNA_MO_DMMA"
ls -la build
