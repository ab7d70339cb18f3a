/* libvcmesh_sampling.so — host-side GraphSampling (connectivity generation).
 *
 * Plain C ABI over vcmeshconv_b200/csrc/graph_sampling.cpp. Replaces the
 * reference's offline tool for the one thing the hot path needs from it: the
 * padded neighbour tables `_pool{i}.npy` / `_unpool{i}.npy`.
 *
 *   reference interface replaced                         entry point here
 *   ---------------------------------------------------  ------------------------
 *   MeshPooler::set_connection_map_from_mesh             vcm_gs_create
 *     (GraphSampling/meshPooler.h:51-63) and
 *     set_must_include_center_lst_from_mesh (:65-80)
 *   MeshCNN::add_pool_layer (meshCNN.h:28-49)            vcm_gs_add_layer
 *   MeshCNN::get_neighborID_lst_lst (meshCNN.h:54-92)    vcm_gs_table_shape / _copy
 *   MeshPooler::_center_lst (meshPooler.h:28; the        vcm_gs_centres
 *     reference writes it corrupt, meshCNN.h:134)
 *
 * Tables are int32, C order, shape [rows, 1 + 2*N]:
 *   row p = [cnt_p, id_0, hop_0, ..., id_{cnt-1}, hop_{cnt-1}, (pad_id, -1) ...]
 * pool table  : rows = centres of the level, ids index the level's INPUT
 *               vertices, pad_id = number of input vertices;
 * unpool table: rows = input vertices of the level, ids index the centres,
 *               pad_id = number of centres.
 * Output is bit-identical to the reference's (tests/test_graph_sampling.py).
 *
 * All functions are thread-compatible (no shared mutable state between
 * handles); errors are returned as negative codes, the message is available
 * from vcm_gs_last_error() on the calling thread. Nothing throws or exits.
 */
#ifndef VCMESH_SAMPLING_H
#define VCMESH_SAMPLING_H

#include <stdint.h>

#if defined(__GNUC__)
#define VCM_GS_API __attribute__((visibility("default")))
#else
#define VCM_GS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vcm_gs vcm_gs;

enum {
    VCM_GS_OK = 0,
    VCM_GS_EINVAL = -1, /* bad argument */
    VCM_GS_EGRAPH = -2  /* level graph has < 3 vertices or is asymmetric (meshPooler.h:82-117) */
};

VCM_GS_API const char *vcm_gs_last_error(void);

/* triangles: [n_tri, 3] 0-based vertex ids; must_include: vertices that have to
 * survive as centres at every level (the reference marks them with colour
 * (0,1,0) in the OBJ). */
VCM_GS_API int vcm_gs_create(const int32_t *triangles, int n_tri, int n_vert, const int32_t *must_include, int n_must,
                  vcm_gs **out);
VCM_GS_API void vcm_gs_destroy(vcm_gs *g);

/* Appends one level. Returns the number of centres (>= 0) or a negative code. */
VCM_GS_API int vcm_gs_add_layer(vcm_gs *g, int stride, int radius_pool, int radius_unpool);
VCM_GS_API int vcm_gs_num_levels(const vcm_gs *g);

VCM_GS_API int vcm_gs_table_shape(const vcm_gs *g, int level, int unpool, int *rows, int *cols);
VCM_GS_API int vcm_gs_table_copy(const vcm_gs *g, int level, int unpool, int32_t *dst);

/* Old ids of the level's centres, ascending. dst may be NULL to query the count. */
VCM_GS_API int vcm_gs_centres(const vcm_gs *g, int level, int32_t *dst, int capacity);

#ifdef __cplusplus
}
#endif
#endif
