// Stand-in for tinyobjloader (not in this image): OBJ meshes are outside the mj_inverse path we
// check; the types exist so the reference's model compiler builds, parsing always reports failure.
#ifndef ORACLE_STUB_TINYOBJ_H_
#define ORACLE_STUB_TINYOBJ_H_
#include <string>
#include <vector>
namespace tinyobj {
struct index_t { int vertex_index; int normal_index; int texcoord_index; };
struct mesh_t {
  std::vector<index_t> indices;
  std::vector<unsigned char> num_face_vertices;
};
struct shape_t { std::string name; mesh_t mesh; };
struct attrib_t { std::vector<float> vertices, normals, texcoords; };
class ObjReader {
 public:
  bool ParseFromString(const std::string&, const std::string&) { return false; }
  bool Valid() const { return false; }
  const attrib_t& GetAttrib() const { return attrib_; }
  const std::vector<shape_t>& GetShapes() const { return shapes_; }
 private:
  attrib_t attrib_;
  std::vector<shape_t> shapes_;
};
}  // namespace tinyobj
#endif
