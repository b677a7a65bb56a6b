// Stand-in for MarchingCubeCpp (not in this image): SDF-plugin meshes are outside the checked path.
#ifndef ORACLE_STUB_MC_H_
#define ORACLE_STUB_MC_H_
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace MC {
typedef float MC_FLOAT;
struct mcVec3f { MC_FLOAT x, y, z; };
struct mcMesh { std::vector<mcVec3f> vertices, normals; std::vector<unsigned> indices; };
inline void marching_cube(MC_FLOAT*, int, int, int, mcMesh&) {
  std::fprintf(stderr, "oracle/_ref: marching cubes is not available in this build\n");
  std::abort();
}
}  // namespace MC
#endif
