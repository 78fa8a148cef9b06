// Host build of the operand-image geometry of the tensor-core route (beta-cores_b200/csrc/bc_umma.cuh) for the CPU tests.
// TEST INFRASTRUCTURE: compiled by tests/test_qsplit_cpu.py with g++; not part of the product.
#include "../../beta-cores_b200/csrc/bc_umma.cuh"
extern "C" {
unsigned q_off(unsigned r, unsigned c) { return bc::q_swizzle_off(r, c); }
int q_slices() { return bc::kQSlices; }
int q_tile_rows() { return bc::kQTileRows; }
int q_chunk() { return bc::kQChunk; }
int q_k() { return bc::kQK; }
}
