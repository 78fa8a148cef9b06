// Host build of the operand-image geometry of the tensor-core route (beta-cores_b200/csrc/bc_umma.cuh) for the CPU tests.
// TEST INFRASTRUCTURE: compiled by tests/test_qsplit_cpu.py with g++; not part of the product.
#include "../../beta-cores_b200/csrc/bc_umma.cuh"
#include "../../beta-cores_b200/csrc/bc_npmean.h"
extern "C" {
unsigned q_off(unsigned r, unsigned c) { return bc::q_swizzle_off(r, c); }
int q_slices() { return bc::kQSlices; }
int q_tile_rows() { return bc::kQTileRows; }
int q_chunk() { return bc::kQChunk; }
int q_k() { return bc::kQK; }
double np_sum_const(double x, int n) { return bc::np_sum_const(x, n); }
double np_score_const(double x, int S, double rsum) { return bc::np_score_const(x, S, rsum); }
}
