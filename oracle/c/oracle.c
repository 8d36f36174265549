/*
 * oracle.c -- CPU restatement (plain C) of the native pieces of the SECOND
 * inference hot path that live in the external `spconv` 1.x package.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path
 * (second.pytorch_b200/) may call into this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED at the spconv boundary: spconv (v1.1..v1.2.1, unpinned by the
 * reference, README.md:104) is not in /root/reference, so there are no golden
 * vectors to check against.  Each function restates the published algorithm and
 * is anchored on the reference's own call sites; the python side cross-checks
 * it against independent known-answer implementations (tests/test_oracle_*.py).
 *
 * Reference anchors (paths relative to /root/reference):
 *   voxelizer   : call site second/data/preprocess.py:303-315, builder
 *                 second/builder/voxel_builder.py:23-32; the one in-tree
 *                 restatement of the index loop is second/utils/simplevis.py:34-50
 *   aligned NMS : second/core/non_max_suppression/nms_gpu.py:10-32,70-126
 *                 (+1 IoU convention, ">" test, 64-wide mask reduce)
 *   nms_cpu     : second/core/non_max_suppression/nms_cpu.py:14-17,34-63
 *   rotated NMS : second/core/non_max_suppression/nms_cpu.py:20-31 (argument
 *                 order), nms_gpu.py:329-401 (in-tree rotated IoU algorithm)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* Voxelizer: spconv points_to_voxel_3d_np<float,3> semantics */
/* ------------------------------------------------------------------------- */
/* pts [P,F] f32; range = lo(xyz),hi(xyz); grid = (gx,gy,gz);
 * voxels [max_voxels,T,F] (zero-filled by caller), coors [max_voxels,3] zyx,
 * num [max_voxels] (zero-filled by caller);
 * scratch = int32 [gz*gy*gx] filled with -1 on entry, restored to -1 on exit.
 * returns voxel_num. */
int orc_points_to_voxel(const float *pts, int P, int F, const float *lo,
                        const float *vs, const int *grid, int T, int max_voxels,
                        float *voxels, int *coors, int *num, int *scratch)
{
    int voxel_num = 0;
    const int gx = grid[0], gy = grid[1], gz = grid[2];
    for (int i = 0; i < P; ++i) {
        int c[3]; /* c[0]=z, c[1]=y, c[2]=x */
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            /* fp32 subtract, fp32 TRUE division, floor (no reciprocal) */
            volatile float d = pts[(size_t)i * F + j] - lo[j];
            volatile float q = d / vs[j];
            int cj = (int)floorf(q);
            if (cj < 0 || cj >= grid[j]) { failed = 1; break; }
            c[2 - j] = cj;
        }
        if (failed) continue;
        size_t cell = ((size_t)c[0] * gy + c[1]) * gx + c[2];
        int vid = scratch[cell];
        if (vid == -1) {
            if (voxel_num >= max_voxels) continue; /* spconv C++: continue */
            vid = voxel_num++;
            scratch[cell] = vid;
            coors[vid * 3 + 0] = c[0];
            coors[vid * 3 + 1] = c[1];
            coors[vid * 3 + 2] = c[2];
        }
        int n = num[vid];
        if (n < T) {
            memcpy(voxels + ((size_t)vid * T + n) * F, pts + (size_t)i * F,
                   sizeof(float) * F);
            num[vid] = n + 1;
        }
    }
    for (int v = 0; v < voxel_num; ++v) {
        size_t cell = ((size_t)coors[v * 3] * gy + coors[v * 3 + 1]) * gx +
                      coors[v * 3 + 2];
        scratch[cell] = -1;
    }
    (void)gz;
    return voxel_num;
}

/* ------------------------------------------------------------------------- */
/* Axis-aligned NMS, "GPU" flavour: spconv.utils.non_max_suppression          */
/* (Fast-R-CNN nms_kernel.cu): boxes already sorted by descending score,      */
/* IoU with +1 on width/height, suppress if IoU > thresh.                     */
/* ------------------------------------------------------------------------- */
static float iou_plus1(const float *a, const float *b)
{
    float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    float width = fmaxf(right - left + 1.f, 0.f);
    float height = fmaxf(bottom - top + 1.f, 0.f);
    float interS = width * height;
    float Sa = (a[2] - a[0] + 1.f) * (a[3] - a[1] + 1.f);
    float Sb = (b[2] - b[0] + 1.f) * (b[3] - b[1] + 1.f);
    return interS / (Sa + Sb - interS);
}

int orc_nms_aligned_sorted(const float *dets /*[N,5]*/, int N, float thresh,
                           int *keep)
{
    unsigned char *removed = (unsigned char *)calloc((size_t)N + 1, 1);
    int nk = 0;
    for (int i = 0; i < N; ++i) {
        if (removed[i]) continue;
        keep[nk++] = i;
        for (int j = i + 1; j < N; ++j) {
            if (removed[j]) continue;
            if (iou_plus1(dets + 5 * i, dets + 5 * j) > thresh) removed[j] = 1;
        }
    }
    free(removed);
    return nk;
}

/* spconv.utils.non_max_suppression_cpu(dets, order, thresh, eps):            */
/* greedy, w = min(x2)-max(x1)+eps, suppress if ovr >= thresh.                */
int orc_nms_cpu(const float *dets /*[N,5]*/, const int *order, int N,
                float thresh, float eps, int *keep)
{
    unsigned char *sup = (unsigned char *)calloc((size_t)N + 1, 1);
    int nk = 0;
    for (int _i = 0; _i < N; ++_i) {
        int i = order[_i];
        if (sup[i]) continue;
        keep[nk++] = i;
        const float *a = dets + 5 * i;
        float area_i = (a[2] - a[0] + eps) * (a[3] - a[1] + eps);
        for (int _j = _i + 1; _j < N; ++_j) {
            int j = order[_j];
            if (sup[j]) continue;
            const float *b = dets + 5 * j;
            float w = fminf(a[2], b[2]) - fmaxf(a[0], b[0]) + eps;
            float h = fminf(a[3], b[3]) - fmaxf(a[1], b[1]) + eps;
            if (w > 0.f && h > 0.f) {
                float area_j = (b[2] - b[0] + eps) * (b[3] - b[1] + eps);
                float inter = w * h;
                float ovr = inter / (area_i + area_j - inter);
                if (ovr >= thresh) sup[j] = 1;
            }
        }
    }
    free(sup);
    return nk;
}

/* ------------------------------------------------------------------------- */
/* Rotated boxes: convex quadrilateral intersection by Sutherland-Hodgman      */
/* clipping, in double.  spconv uses boost::geometry intersection/union on     */
/* the 4-corner polygons; for convex quads area(A∩B) is what S-H computes and  */
/* area(A∪B) = area(A)+area(B)-area(A∩B).                                      */
/* ------------------------------------------------------------------------- */
#define REAL double
#define FN(name) name
#include "quad_clip.inc"
#undef REAL
#undef FN
#define REAL float
#define FN(name) name##_f32
#include "quad_clip.inc"
#undef REAL
#undef FN

/* spconv.utils.rotate_non_max_suppression_cpu(corners, order, standup_iou, thresh):
 * greedy in `order`; skip pair if standup_iou <= 0; suppress if inter/union >= thresh.
 * iou_out (optional, [N,N] f64, pre-filled with -1): records every IoU evaluated, so
 * tests can exclude near-threshold pairs. */
int orc_rotate_nms(const float *corners /*[N,4,2]*/, const int *order,
                   const float *standup_iou /*[N,N]*/, int N, float thresh,
                   int *keep, double *iou_out)
{
    unsigned char *sup = (unsigned char *)calloc((size_t)N + 1, 1);
    int nk = 0;
    for (int _i = 0; _i < N; ++_i) {
        int i = order[_i];
        if (sup[i]) continue;
        keep[nk++] = i;
        for (int _j = _i + 1; _j < N; ++_j) {
            int j = order[_j];
            if (sup[j]) continue;
            if (standup_iou[(size_t)i * N + j] <= 0.f) continue;
            double sa, sb;
            double inter = orc_quad_intersection(corners + 8 * i, corners + 8 * j,
                                                 &sa, &sb);
            if (inter <= 0.0) continue; /* empty intersection: no suppression */
            double uni = sa + sb - inter;
            double ov = inter / uni;
            if (iou_out) iou_out[(size_t)i * N + j] = ov;
            if (ov >= (double)thresh) sup[j] = 1;
        }
    }
    free(sup);
    return nk;
}

/* the same greedy loop with the polygon clip in fp32 (see quad_clip.inc) */
int orc_rotate_nms_f32(const float *corners /*[N,4,2]*/, const int *order,
                       const float *standup_iou /*[N,N]*/, int N, float thresh,
                       int *keep, double *iou_out)
{
    unsigned char *sup = (unsigned char *)calloc((size_t)N + 1, 1);
    int nk = 0;
    for (int _i = 0; _i < N; ++_i) {
        int i = order[_i];
        if (sup[i]) continue;
        keep[nk++] = i;
        for (int _j = _i + 1; _j < N; ++_j) {
            int j = order[_j];
            if (sup[j]) continue;
            if (standup_iou[(size_t)i * N + j] <= 0.f) continue;
            float sa, sb;
            float inter = orc_quad_intersection_f32(corners + 8 * i, corners + 8 * j, &sa, &sb);
            if (inter <= 0.f) continue;
            float uni = sa + sb - inter;
            float ov = inter / uni;
            if (iou_out) iou_out[(size_t)i * N + j] = (double)ov;
            if (ov >= thresh) sup[j] = 1;
        }
    }
    free(sup);
    return nk;
}

/* spconv.utils.rbbox_iou / rbbox_intersection(corners[N,4,2], qcorners[K,4,2],
 * standup_iou[N,K], standup_thresh) -> [N,K].  mode 0: IoU, mode 1: intersection AREA.
 * In-tree anchor for mode 1: second/utils/eval.py:174-175 uses rinter_cc (box_np_ops.py:23-34) interchangeably with
 * rotate_iou_gpu_eval(..., criterion=2), whose device function returns area_inter (nms_gpu.py:553-566). */
void orc_rbbox_iou(const float *corners, const float *qcorners,
                   const float *standup_iou, int N, int K, float standup_thresh,
                   int mode, float *out)
{
    for (int k = 0; k < K; ++k)
        for (int n = 0; n < N; ++n) {
            float r = 0.f;
            if (standup_iou[(size_t)n * K + k] > standup_thresh) {
                double sa, sb;
                double inter = orc_quad_intersection(corners + 8 * n,
                                                     qcorners + 8 * k, &sa, &sb);
                if (mode == 0) r = (float)(inter / (sa + sb - inter));
                else r = (float)inter;
            }
            out[(size_t)n * K + k] = r;
        }
}

int orc_version(void) { return 1; }
