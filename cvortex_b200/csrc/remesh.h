// remesh.h -- internal interface of the grid redistribution: the node set both the
// device stage (remesh_device.cu) and the host stage (remesh.cpp) produce, and the grid
// placement they share.  No CUDA types.
#pragma once
#include <cstdint>
#include <vector>
#include "remesh_math.h"

namespace cvtx {
namespace remesh {

// Grid nodes that received vorticity, in ascending Morton order of their indices, with
// the summed strength (3 components per node in 3-D, 1 in 2-D).
struct NodeSet {
	std::vector<uint64_t> code;
	std::vector<float> strength;
};

// Where the grid sits for a given particle set (reference src/P3D.cpp:538-550,
// src/P2D.cpp:312-323): a node coincides with the mean position, and the origin is
// pushed at least `half` + 5 cells below the lowest particle.  `rows` are the gathered
// particle structs (row_floats floats each, coordinates first) -- gathered here, in the same
// pass, when `particles` (the caller's array of pointers) is not null; `kind` is -1 and `half` the
// caller's (int)roundf(radius) for a user-defined interpolant.  max_index receives the
// largest node index any stencil can touch.
Grid place_grid(int dim, int kind, int half, float h, const void *const *particles, float *rows, long n, int row_floats,
                uint32_t *max_index);

// Bits of Morton code needed for indices <= max_index; -1 when the grid is larger than
// the codes can hold (2^21 nodes per axis in 3-D, 2^31 in 2-D).
int code_bits(int dim, uint32_t max_index);

// Device stage: particles (array of pointers to cvtx_P3D / cvtx_P2D) -> node set, on
// `device`.  Returns a cvtx_b200_status; fills grid and nodes on success.
int device_nodes(int device, int dim, int kind, float h, const void *const *particles, long n, Grid *grid, NodeSet *nodes);

}  // namespace remesh
}  // namespace cvtx
