// TEST INFRASTRUCTURE — not shipped, not on the product path.
//
// Thin driver around the *unmodified* reference GraphSampling headers, compiled
// from where they lie under /root/reference (see oracle/Makefile). It exists
// because the reference's own main() is unusable:
//   * Mesh::loadmesh_obj (GraphSampling/meshLoader.h:45-54) has no return on the
//     success path (UB), so we call Mesh::LoadOBJ_withcolor (meshLoader.h:221)
//     directly;
//   * the layer preset is chosen by editing main.cpp:311, so this driver takes
//     the (stride, pool_radius, unpool_radius) list on the command line.
//
// Usage: graphsampling_ref mesh.obj out_prefix s0,rp0,ru0 s1,rp1,ru1 ...
// Writes out_prefix + "_pool{i}.npy" / "_unpool{i}.npy" exactly as
// MeshCNN::save_pool_and_unpool_neighbor_info_to_npz (meshCNN.h:99-139) does.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "meshCNN.h"

int main(int argc, char **argv) {
    if (argc < 4) {
        fprintf(stderr, "usage: %s mesh.obj out_prefix s,rp,ru [s,rp,ru ...]\n", argv[0]);
        return 2;
    }
    Mesh mesh;
    if (!mesh.LoadOBJ_withcolor(argv[1], mesh.points, mesh.colors, mesh.triangles)) {
        fprintf(stderr, "cannot load %s\n", argv[1]);
        return 3;
    }
    MeshCNN cnn(mesh);
    for (int a = 3; a < argc; ++a) {
        int s = 0, rp = 0, ru = 0;
        if (sscanf(argv[a], "%d,%d,%d", &s, &rp, &ru) != 3) {
            fprintf(stderr, "bad layer spec %s\n", argv[a]);
            return 2;
        }
        cnn.add_pool_layer(s, rp, ru);
        // meshCNN.h:113 divides by the level's vertex count; a level that fell
        // below 3 vertices leaves empty maps (meshPooler.h:82-89) -> SIGFPE.
        if (cnn._meshPoolers.back()._center_center_map.size() < 1) {
            fprintf(stderr, "level %d collapsed; refusing to save\n", a - 3);
            return 4;
        }
    }
    cnn.save_pool_and_unpool_neighbor_info_to_npz(std::string(argv[2]));
    return 0;
}
