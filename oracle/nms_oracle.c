/* C restatement of the reference's class-aware greedy NMS -- TEST INFRASTRUCTURE ONLY.
 *
 * Follows code/utils.py:150-191 (non_max_suppression) and code/utils.py:38-84 (calc_iou) of
 * GabeTsai/YOLO-For-Turbines op for op in fp32 (compile with -ffp-contract=off so that no
 * mul+add is fused): it exists so that parity tests at 10^4..10^5 boxes per image finish in
 * seconds; oracle/yolo_oracle.py::non_max_suppression is the literal (slow) restatement and
 * tests/test_oracle_golden.py pins both against outputs of the reference itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library.
 */
#include <stdint.h>
#include <stdlib.h>

static inline float nanmax(float a, float b) { return (a > b || a != a) ? a : b; } /* torch.max */
static inline float nanmin(float a, float b) { return (a < b || a != a) ? a : b; } /* torch.min */
static inline float clamp0(float d) { return d < 0.0f ? 0.0f : d; }               /* torch.clamp(min=0) */

typedef struct { float x1, y1, x2, y2, area, score, cls; int idx; } obox;

static float iou(const obox* a, const obox* b) {
  float xa = nanmax(a->x1, b->x1), ya = nanmax(a->y1, b->y1);       /* utils.py:70-71 */
  float xb = nanmin(a->x2, b->x2), yb = nanmin(a->y2, b->y2);       /* utils.py:72-73 */
  float iw = clamp0(xb - xa), ih = clamp0(yb - ya);                 /* utils.py:75-76 */
  float inter = iw * ih;                                            /* utils.py:77 */
  float uni = a->area + b->area;                                    /* utils.py:81 */
  uni = uni - inter;
  uni = uni + 1e-6f;                                                /* utils.py:83 */
  return inter / uni;
}

static int cmp_desc(const void* pa, const void* pb) {
  const obox *a = (const obox*)pa, *b = (const obox*)pb;
  if (a->score > b->score) return -1;                               /* utils.py:166 reverse=True */
  if (a->score < b->score) return 1;
  return (a->idx > b->idx) - (a->idx < b->idx);                     /* stable: ties keep input order */
}

/* boxes: n rows [x,y,w,h,score,cls]; center != 0 -> "center" format, else top-left xywh.
 * keep_out receives the row indices of the survivors in the reference's output order.
 * Returns the number kept, or -1 on allocation failure. */
int oracle_nms(const float* boxes, int n, float iou_thr, double obj_thr, int center, int32_t* keep_out) {
  obox* c = (obox*)malloc(sizeof(obox) * (size_t)(n > 0 ? n : 1));
  uint8_t* gone;
  int m = 0, kept = 0;
  if (!c) return -1;
  for (int i = 0; i < n; ++i) {
    const float* r = boxes + (size_t)i * 6;
    if (!((double)r[4] > obj_thr)) continue;                        /* utils.py:165 (Python floats) */
    obox* o = &c[m++];
    if (center) {                                                   /* utils.py:60,63 */
      o->x1 = r[0] - r[2] / 2.0f;
      o->y1 = r[1] - r[3] / 2.0f;
    } else {                                                        /* utils.py:66-67 */
      o->x1 = r[0];
      o->y1 = r[1];
    }
    o->x2 = o->x1 + r[2];
    o->y2 = o->y1 + r[3];
    o->area = r[2] * r[3];                                          /* utils.py:79-80 */
    o->score = r[4];
    o->cls = r[5];
    o->idx = i;
  }
  qsort(c, (size_t)m, sizeof(obox), cmp_desc);
  gone = (uint8_t*)calloc((size_t)(m > 0 ? m : 1), 1);
  if (!gone) { free(c); return -1; }
  for (int i = 0; i < m; ++i) {                                     /* utils.py:170-187 */
    if (gone[i]) continue;
    keep_out[kept++] = c[i].idx;
    for (int j = i + 1; j < m; ++j) {
      if (gone[j]) continue;
      if (c[j].cls != c[i].cls) continue;                           /* class_mask keeps it */
      if (iou(&c[i], &c[j]) < iou_thr) continue;                    /* iou_mask keeps it   */
      gone[j] = 1;                                                  /* NaN IoU lands here  */
    }
  }
  free(gone);
  free(c);
  return kept;
}
