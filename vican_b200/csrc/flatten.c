/* Host-side flatten of the reference's detection dictionary (CPython extension `vican_b200._vb_flatten`).
 *
 * The reference walks `src_edges` twice in Python (vican/bipgo.py:203-221 for the rotation stage, :420-471 for
 * the translation stage), calling `edge_filter`, `noise_model_r`, `noise_model_t` and `pose.R()` / `pose.t()` per
 * detection.  The callables are the caller's Python and have to run on the host; everything around them (key
 * coding, copying 9 + 3 numbers per pose, collecting the weights) is done here in one pass over `items()` instead
 * of seven list comprehensions + np.array over two million small arrays.
 *
 *   flatten(src_edges, edge_filter, noise_model_r, noise_model_t, as_f64, R_out, t_out, kr_out, kt_out,
 *           cam_code, tm_code) -> (n_kept, cam_keys, tm_keys, r_format, kr_first)
 *
 * Outputs are caller-allocated writable C-contiguous buffers sized for len(src_edges): float64 R_out[n][9],
 * t_out[n][3], kr_out[n], kt_out[n]; int32 cam_code[n], tm_code[n].  cam_keys / tm_keys are the DISTINCT first /
 * second key components in first-seen order, the code arrays index into them.  r_format is the struct format of
 * the pose rotation arrays ('d', 'f', ...; detections mixing formats raise ValueError like the Python flatten),
 * kr_first the first object noise_model_r returned (its TYPE decides numpy's float32 product, bipgo.py:212).
 * `as_f64` is a Python callable (np.ascontiguousarray(x, dtype=float64)) used for poses whose arrays are neither
 * float64 nor float32 buffers.  Callables see exactly the kept detections: edge_filter once per detection, the
 * noise models once per kept detection.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static PyObject *s_pose, *s_R, *s_t;

/* copy `count` numbers of a 1-D/2-D float buffer (format 'd' or 'f') into dst; returns 0 on success,
   1 if the object has to go through the converter, -1 on error */
static int copy_small(PyObject *obj, double *dst, int rows, int cols, char *fmt_out)
{
    Py_buffer v;
    if (PyObject_GetBuffer(obj, &v, PyBUF_STRIDED_RO | PyBUF_FORMAT) != 0) {
        PyErr_Clear();
        return 1;
    }
    int rc = 1;
    const char *f = v.format ? v.format : "B";
    if (f[0] == '=' || f[0] == '<' || f[0] == '@') f++;
    char c = f[0];
    int is_d = (c == 'd' && f[1] == 0 && v.itemsize == 8), is_f = (c == 'f' && f[1] == 0 && v.itemsize == 4);
    Py_ssize_t total = 1;
    for (int i = 0; i < v.ndim; ++i) total *= v.shape[i];
    if (fmt_out) *fmt_out = (f[1] == 0) ? c : '?';
    if ((is_d || is_f) && total == (Py_ssize_t)rows * cols) {
        const char *base = (const char *)v.buf;
        if (v.ndim == 2 && v.shape[0] == rows && v.shape[1] == cols) {
            for (int i = 0; i < rows; ++i)
                for (int j = 0; j < cols; ++j) {
                    const char *p = base + i * v.strides[0] + j * v.strides[1];
                    dst[i * cols + j] = is_d ? *(const double *)p : (double)*(const float *)p;
                }
            rc = 0;
        } else if (v.ndim == 1) {
            for (int i = 0; i < rows * cols; ++i) {
                const char *p = base + i * v.strides[0];
                dst[i] = is_d ? *(const double *)p : (double)*(const float *)p;
            }
            rc = 0;
        }
    }
    PyBuffer_Release(&v);
    return rc;
}

static int copy_pose_part(PyObject *obj, PyObject *as_f64, double *dst, int rows, int cols, char *fmt_out)
{
    int rc = copy_small(obj, dst, rows, cols, fmt_out);
    if (rc <= 0) return rc;
    /* other dtypes / shapes ((3,1) columns, lists, ...): numpy converts, flattened */
    PyObject *conv = PyObject_CallOneArg(as_f64, obj);
    if (!conv) return -1;
    Py_buffer v;
    if (PyObject_GetBuffer(conv, &v, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) { Py_DECREF(conv); return -1; }
    if (v.len != (Py_ssize_t)sizeof(double) * rows * cols) {
        PyBuffer_Release(&v); Py_DECREF(conv);
        PyErr_Format(PyExc_ValueError, "pose array has %zd bytes as float64, expected %d x %d numbers", v.len, rows, cols);
        return -1;
    }
    memcpy(dst, v.buf, sizeof(double) * rows * cols);
    PyBuffer_Release(&v); Py_DECREF(conv);
    return 0;
}

static int code_of(PyObject *table, PyObject *list, PyObject *key, int32_t *out)
{
    PyObject *hit = PyDict_GetItemWithError(table, key);                /* borrowed */
    if (hit) { *out = (int32_t)PyLong_AsLong(hit); return 0; }
    if (PyErr_Occurred()) return -1;
    Py_ssize_t n = PyList_GET_SIZE(list);
    PyObject *num = PyLong_FromSsize_t(n);
    if (!num) return -1;
    int rc = PyDict_SetItem(table, key, num);
    Py_DECREF(num);
    if (rc != 0 || PyList_Append(list, key) != 0) return -1;
    *out = (int32_t)n;
    return 0;
}

typedef struct { Py_buffer v; int held; } outbuf;

static int get_out(PyObject *o, outbuf *b, Py_ssize_t need_bytes, const char *name)
{
    if (PyObject_GetBuffer(o, &b->v, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) return -1;
    b->held = 1;
    if (b->v.len < need_bytes) {
        PyErr_Format(PyExc_ValueError, "%s: buffer of %zd bytes, need %zd", name, b->v.len, need_bytes);
        return -1;
    }
    return 0;
}

static PyObject *flatten(PyObject *self, PyObject *args)
{
    PyObject *edges, *filt, *nr, *nt, *as_f64, *oR, *ot, *okr, *okt, *occ, *otc;
    if (!PyArg_ParseTuple(args, "OOOOOOOOOOO", &edges, &filt, &nr, &nt, &as_f64, &oR, &ot, &okr, &okt, &occ, &otc))
        return NULL;
    Py_ssize_t n = PyObject_Length(edges);
    if (n < 0) return NULL;
    outbuf b[6];
    memset(b, 0, sizeof(b));
    PyObject *items = NULL, *it = NULL, *cam_tab = NULL, *tm_tab = NULL, *cam_keys = NULL, *tm_keys = NULL;
    PyObject *kr_first = NULL, *result = NULL, *item = NULL;
    if (get_out(oR, &b[0], n * 72, "R_out") || get_out(ot, &b[1], n * 24, "t_out") ||
        get_out(okr, &b[2], n * 8, "kr_out") || get_out(okt, &b[3], n * 8, "kt_out") ||
        get_out(occ, &b[4], n * 4, "cam_code") || get_out(otc, &b[5], n * 4, "tm_code"))
        goto done;
    double *R = (double *)b[0].v.buf, *t = (double *)b[1].v.buf, *kr = (double *)b[2].v.buf, *kt = (double *)b[3].v.buf;
    int32_t *cc = (int32_t *)b[4].v.buf, *tc = (int32_t *)b[5].v.buf;
    items = PyObject_CallMethod(edges, "items", NULL);
    if (!items) goto done;
    it = PyObject_GetIter(items);
    cam_tab = PyDict_New(); tm_tab = PyDict_New(); cam_keys = PyList_New(0); tm_keys = PyList_New(0);
    if (!it || !cam_tab || !tm_tab || !cam_keys || !tm_keys) goto done;
    Py_ssize_t k = 0;
    char r_fmt = 0;
    while ((item = PyIter_Next(it)) != NULL) {
        if (!PyTuple_Check(item) || PyTuple_GET_SIZE(item) != 2) {
            PyErr_SetString(PyExc_TypeError, "src_edges.items() must yield (key, value) pairs"); goto done;
        }
        PyObject *key = PyTuple_GET_ITEM(item, 0), *val = PyTuple_GET_ITEM(item, 1);
        PyObject *keep = PyObject_CallOneArg(filt, val);                 /* bipgo.py:204, :423 */
        if (!keep) goto done;
        int truth = PyObject_IsTrue(keep);
        Py_DECREF(keep);
        if (truth < 0) goto done;
        if (!truth) { Py_CLEAR(item); continue; }
        if (k >= n) { PyErr_SetString(PyExc_RuntimeError, "src_edges changed size during the flatten"); goto done; }
        PyObject *k0 = PySequence_GetItem(key, 0);
        PyObject *k1 = k0 ? PySequence_GetItem(key, 1) : NULL;
        int rc = (k0 && k1) ? (code_of(cam_tab, cam_keys, k0, &cc[k]) || code_of(tm_tab, tm_keys, k1, &tc[k])) : -1;
        Py_XDECREF(k0); Py_XDECREF(k1);
        if (rc) goto done;
        PyObject *pose = PyObject_GetItem(val, s_pose);
        if (!pose) goto done;
        PyObject *Ro = PyObject_CallMethodNoArgs(pose, s_R);
        PyObject *to = Ro ? PyObject_CallMethodNoArgs(pose, s_t) : NULL;
        Py_DECREF(pose);
        char fmt = '?';
        rc = (Ro && to) ? (copy_pose_part(Ro, as_f64, R + 9 * k, 3, 3, &fmt) || copy_pose_part(to, as_f64, t + 3 * k, 3, 1, NULL)) : -1;
        Py_XDECREF(Ro); Py_XDECREF(to);
        if (rc) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "bad pose arrays"); goto done; }
        if (k == 0) r_fmt = fmt;
        else if (fmt != r_fmt) {
            PyErr_SetString(PyExc_ValueError, "detections mix float32 and float64 pose arrays; convert them to one dtype");
            goto done;
        }
        PyObject *w = PyObject_CallOneArg(nr, val);                      /* bipgo.py:212 */
        if (!w) goto done;
        kr[k] = PyFloat_AsDouble(w);
        if (k == 0) kr_first = w; else Py_DECREF(w);
        if (kr[k] == -1.0 && PyErr_Occurred()) goto done;
        w = PyObject_CallOneArg(nt, val);                                /* bipgo.py:449 */
        if (!w) goto done;
        kt[k] = PyFloat_AsDouble(w);
        Py_DECREF(w);
        if (kt[k] == -1.0 && PyErr_Occurred()) goto done;
        ++k;
        Py_CLEAR(item);
    }
    if (PyErr_Occurred()) goto done;
    if (!kr_first) { kr_first = Py_None; Py_INCREF(kr_first); }
    {
        char fs[2] = { r_fmt ? r_fmt : 'd', 0 };
        result = Py_BuildValue("nOOsO", k, cam_keys, tm_keys, fs, kr_first);
    }
done:
    Py_XDECREF(item); Py_XDECREF(items); Py_XDECREF(it); Py_XDECREF(cam_tab); Py_XDECREF(tm_tab);
    Py_XDECREF(cam_keys); Py_XDECREF(tm_keys); Py_XDECREF(kr_first);
    for (int i = 0; i < 6; ++i) if (b[i].held) PyBuffer_Release(&b[i].v);
    return result;
}

/* poses(vals, as_f64, R_out, t_out) -> r_format
 * Copies value["pose"].R() / .t() of every detection of the LIST `vals` into float64 R_out[n][9], t_out[n][3] (what
 * object_bipartite_se3sync needs to invert all poses as one device batch, bipgo.py:526-531), float32 arrays converted
 * exactly (what np.stack's upcast does).  Returns the common struct format of the R arrays, '?' if they differ. */
static PyObject *poses(PyObject *self, PyObject *args)
{
    PyObject *vals, *as_f64, *oR, *ot;
    if (!PyArg_ParseTuple(args, "O!OOO", &PyList_Type, &vals, &as_f64, &oR, &ot)) return NULL;
    const Py_ssize_t n = PyList_GET_SIZE(vals);
    outbuf b[2];
    memset(b, 0, sizeof(b));
    PyObject *result = NULL;
    char r_fmt = 0;
    if (get_out(oR, &b[0], n * 72, "R_out") || get_out(ot, &b[1], n * 24, "t_out")) goto done;
    double *R = (double *)b[0].v.buf, *t = (double *)b[1].v.buf;
    for (Py_ssize_t k = 0; k < n; ++k) {
        if (k >= PyList_GET_SIZE(vals)) { PyErr_SetString(PyExc_RuntimeError, "list changed size"); goto done; }
        PyObject *pose = PyObject_GetItem(PyList_GET_ITEM(vals, k), s_pose);
        if (!pose) goto done;
        PyObject *Ro = PyObject_CallMethodNoArgs(pose, s_R);
        PyObject *to = Ro ? PyObject_CallMethodNoArgs(pose, s_t) : NULL;
        Py_DECREF(pose);
        char fmt = '?';
        int rc = (Ro && to) ? (copy_pose_part(Ro, as_f64, R + 9 * k, 3, 3, &fmt) || copy_pose_part(to, as_f64, t + 3 * k, 3, 1, NULL)) : -1;
        Py_XDECREF(Ro); Py_XDECREF(to);
        if (rc) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_ValueError, "bad pose arrays"); goto done; }
        if (k == 0) r_fmt = fmt;
        else if (fmt != r_fmt) r_fmt = '?';          /* mixed dtypes: values are converted like np.stack's upcast */
    }
    {
        char fs[2] = { r_fmt ? r_fmt : 'd', 0 };
        result = PyUnicode_FromString(fs);
    }
done:
    for (int i = 0; i < 2; ++i) if (b[i].held) PyBuffer_Release(&b[i].v);
    return result;
}

/* rekey(keys, vals, inv_poses, root) -> dict
 * The re-keyed detection dictionary of object_bipartite_se3sync (bipgo.py:526-531): for key (c, "t_m") and value v,
 *     out[(m, t + "_" + root)] = {"pose": inv_poses[i], "corners": v["corners"], "reprojected_err": v["reprojected_err"],
 *                                 "im_filename": v["im_filename"]}
 * in the order of the lists (a repeated new key keeps the later detection, like the reference's dict assignment).
 * Missing fields raise KeyError, a second key component that is not "t_m" raises ValueError, as the Python loop does. */
static PyObject *rekey(PyObject *self, PyObject *args)
{
    PyObject *keys, *vals, *inv, *root;
    if (!PyArg_ParseTuple(args, "O!O!O!U", &PyList_Type, &keys, &PyList_Type, &vals, &PyList_Type, &inv, &root)) return NULL;
    const Py_ssize_t n = PyList_GET_SIZE(keys);
    if (PyList_GET_SIZE(vals) != n || PyList_GET_SIZE(inv) != n) {
        PyErr_SetString(PyExc_ValueError, "keys, vals and inv_poses must have the same length");
        return NULL;
    }
    PyObject *out = PyDict_New(), *sep = PyUnicode_FromString("_");
    PyObject *s_corners = PyUnicode_InternFromString("corners"), *s_err = PyUnicode_InternFromString("reprojected_err");
    PyObject *s_im = PyUnicode_InternFromString("im_filename");
    PyObject *tail = sep ? PyUnicode_Concat(sep, root) : NULL;          /* "_" + root */
    int ok = out && sep && s_corners && s_err && s_im && tail;
    for (Py_ssize_t i = 0; ok && i < n; ++i) {
        PyObject *k1 = PySequence_GetItem(PyList_GET_ITEM(keys, i), 1);
        PyObject *parts = k1 ? PyUnicode_Split(k1, sep, -1) : NULL;
        Py_XDECREF(k1);
        if (!parts) { ok = 0; break; }
        if (PyList_GET_SIZE(parts) != 2) {
            PyErr_Format(PyExc_ValueError, "%s values to unpack (expected 2)", PyList_GET_SIZE(parts) < 2 ? "not enough" : "too many");
            Py_DECREF(parts); ok = 0; break;
        }
        PyObject *tkey = PyUnicode_Concat(PyList_GET_ITEM(parts, 0), tail);
        PyObject *nk = tkey ? PyTuple_Pack(2, PyList_GET_ITEM(parts, 1), tkey) : NULL;
        Py_XDECREF(tkey); Py_DECREF(parts);
        PyObject *v = PyList_GET_ITEM(vals, i);
        PyObject *c = nk ? PyObject_GetItem(v, s_corners) : NULL;
        PyObject *e = c ? PyObject_GetItem(v, s_err) : NULL;
        PyObject *f = e ? PyObject_GetItem(v, s_im) : NULL;
        PyObject *d = f ? PyDict_New() : NULL;
        if (!d || PyDict_SetItem(d, s_pose, PyList_GET_ITEM(inv, i)) || PyDict_SetItem(d, s_corners, c) ||
            PyDict_SetItem(d, s_err, e) || PyDict_SetItem(d, s_im, f) || PyDict_SetItem(out, nk, d))
            ok = 0;
        Py_XDECREF(nk); Py_XDECREF(c); Py_XDECREF(e); Py_XDECREF(f); Py_XDECREF(d);
    }
    Py_XDECREF(sep); Py_XDECREF(s_corners); Py_XDECREF(s_err); Py_XDECREF(s_im); Py_XDECREF(tail);
    if (!ok) { Py_XDECREF(out); return NULL; }
    return out;
}

static PyMethodDef methods[] = {
    {"flatten", flatten, METH_VARARGS, "one-pass flatten of a detection dictionary (see csrc/flatten.c)"},
    {"poses", poses, METH_VARARGS, "stack pose.R() / pose.t() of a list of detections into float64 arrays"},
    {"rekey", rekey, METH_VARARGS, "the re-keyed detection dictionary of object_bipartite_se3sync"},
    {NULL, NULL, 0, NULL}
};

static struct PyModuleDef moddef = { PyModuleDef_HEAD_INIT, "_vb_flatten", NULL, -1, methods };

PyMODINIT_FUNC PyInit__vb_flatten(void)
{
    s_pose = PyUnicode_InternFromString("pose");
    s_R = PyUnicode_InternFromString("R");
    s_t = PyUnicode_InternFromString("t");
    if (!s_pose || !s_R || !s_t) return NULL;
    return PyModule_Create(&moddef);
}
