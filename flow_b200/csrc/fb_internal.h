// Internal host-side structures shared by fb_topology.cpp (pure host) and the CUDA units.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/flowb200.h"

struct fb_device_state;  // defined in fb_device.cuh (CUDA units only)

struct fb_comm;  // NCCL communicator state (fb_comm.cu)

struct fb_ctx {
  int device = -1;
  std::string err;
  fb_device_state *dev = nullptr;  // null for host-only contexts
  int64_t launches = 0;
  int64_t halo_calls = 0, allreduce_calls = 0;  // exchanges / reductions enqueued by partitioned runs
  fb_comm *comm = nullptr;  // null: single rank
};

struct fb_mesh {
  fb_ctx *ctx = nullptr;
  int dim = 0;
  int64_t nv = 0, nc = 0, ne = 0;
  std::vector<double> xyz;          // nv*dim
  std::vector<int32_t> cells;       // nc*(dim+1), ascending per cell
  std::vector<int32_t> edges;       // ne*2, lexicographic
  std::vector<int32_t> cell_edges;  // nc*NE, UFC local edge order
  std::vector<int32_t> bf_cell, bf_local;
  std::vector<uint8_t> bvert, bedge;
  std::vector<int32_t> cell_order;  // device storage order of the cells (Morton order of the centroids), built lazily
  void *dev = nullptr;  // DeviceMesh*
};

struct fb_space {
  fb_mesh *mesh = nullptr;
  int degree = 1, ncomp = 1, nl = 0;
  int64_t nnodes = 0;
  std::vector<int32_t> cell_nodes;  // nc*nl
  std::vector<double> coords;       // nnodes*dim
  std::vector<uint8_t> bnode;       // nnodes
  std::vector<int64_t> indptr;      // node-level pattern, built lazily
  std::vector<int32_t> indices;
  // distributed runs: nodes [0, n_owned) are owned by this rank, [n_owned, nnodes) are ghosts
  // grouped by owner; halo plan per neighbour rank
  int64_t n_owned = 0;
  std::vector<int32_t> halo_ranks;
  std::vector<int64_t> halo_send_ptr, halo_recv_ptr;  // nneigh + 1 (recv offsets relative to n_owned)
  std::vector<int32_t> halo_send_nodes;
  void *dev = nullptr;  // DeviceSpace*
};

int fb_fail(fb_ctx *ctx, int status, const std::string &msg);
int fb_space_build_pattern(fb_space *s);
// new position -> mesh cell: Morton order of the cell centroids (identity with FB_CELL_ORDER=0)
const std::vector<int32_t> &fb_mesh_cell_order(fb_mesh *m);

static inline int fb_num_local_edges(int dim) { return dim == 2 ? 3 : 6; }
