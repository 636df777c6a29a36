// mrh_mesh.cu — extractMesh (placeholder until the marching-cubes milestone lands)
#include "mrh_host.h"
using namespace mrh;
extern "C" {
int mrh_extract_mesh(mrh_map* m, const char* path) {
  (void) m, (void) path;
  return fail("mrh_extract_mesh: marching cubes not built yet");
}
int mrh_get_mesh(mrh_map* m, const double** v, const int32_t** f, const double** c, size_t* nv, size_t* nf) {
  (void) m, (void) v, (void) f, (void) c, (void) nv, (void) nf;
  return fail("mrh_get_mesh: marching cubes not built yet");
}
int mrh_get_triangles(mrh_map* m, const float** t, size_t* n) {
  (void) m, (void) t, (void) n;
  return fail("mrh_get_triangles: marching cubes not built yet");
}
}
