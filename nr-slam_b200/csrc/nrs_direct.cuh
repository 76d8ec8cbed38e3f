// nrs_direct.cuh — launch interface of the exact-solve tracking engine (nrs_direct.cu).
#pragma once
#include <cuda_runtime.h>

#include "nrs_direct_core.cuh"
#include "nrs_engine.cuh"

namespace nrs {

struct DirectParams {
  Params P;            // the staged tracking problem (rows in elimination order)
  direct::Plan pl;     // elimination tree + factor storage
  const int* inc_pos;  // [2P] front position of the other endpoint of every incidence (nrs_direct_plan.h)
  double* cpl;         // [18V] pose coupling blocks of the current linearisation
  double* hpp_part;    // [G][28] per-CTA partial of H_pp / b_p
  double* hpp;         // [28]   their sum (published by CTA 0 for the root front)
  double* dslots;      // [2][G][2] grid-reduction slots
  double* dpose;       // [8]    pose part of the last solution
  int scratch_z;       // shared-memory doubles of the backward substitution scratch
  int max_nv;          // most own vertices of any front
  int max_rows;        // most panel rows of any team member
  unsigned long long* tbar;  // [2 (n_nodes + 2)] per-node barrier counters: 2t after stage AB, 2t + 1 children done
  long long* plev;     // [G][32] per-level cycle counters of every CTA (diagnostics, may be null)
};

size_t direct_smem_bytes(int max_path, int scratch_z, int max_nv, int max_rows, size_t panel_doubles);
int direct_block_threads();
int direct_max_grid(size_t smem);  // co-resident CTAs of the kernel with this much dynamic shared memory
int launch_direct(const DirectParams& q, int grid, size_t smem, cudaStream_t stream);

}  // namespace nrs
