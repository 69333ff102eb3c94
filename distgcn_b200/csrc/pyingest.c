/* _pyingest: walk a Python sequence of scipy sparse matrices (or of numpy arrays) in C and hand back tables of raw
 * pointers for the C-ABI ingest entry points (dg_pack_graphs_host / dg_solve_graphs_host, include/distgcn_b200.h).
 *
 * The reference keeps one scipy matrix per graph (mwis_dqn_call.py:198; .mat files load as CSC with int32 indptr /
 * indices and float64 data).  Getting 2-3 buffer addresses per graph from Python costs ~1 us each; here the loop runs
 * in C (~0.1 us per graph), so a 500-graph call spends its time in the packer and on the GPU, not in the interpreter.
 *
 *   collect(seq, want_data) -> (indptr_tab, indices_tab, data_tab | None, n_rows, keepalive, n_nodes)
 *       *_tab: bytes objects holding n pointers (uintptr_t) / n int32; keepalive: list of the attribute objects whose
 *       buffers the pointers refer to (hold it for the duration of the native call).  Raises TypeError when an item
 *       has no int32 C-contiguous indptr / indices (the caller then normalises that input in Python).
 *   pointers(seq, itemsize) -> (tab, lengths, keepalive)
 *       the same for a sequence of 1-D contiguous arrays (per-graph weight vectors).
 * Nothing here computes on the data.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

static PyObject *s_indptr, *s_indices, *s_data;

/* address of a C-contiguous buffer with the given item size; returns 0 and sets *addr, *count */
static int buffer_of(PyObject *obj, Py_ssize_t itemsize, char kind, uintptr_t *addr, Py_ssize_t *count) {
    Py_buffer view;
    if (PyObject_GetBuffer(obj, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) return -1;
    int ok = view.itemsize == itemsize;
    if (ok && view.format) {
        const char *f = view.format;
        while (*f == '<' || *f == '=' || *f == '@') ++f;
        if (kind == 'i') ok = (*f == 'i' || *f == 'l' || *f == 'q' || *f == 'n') && view.itemsize == itemsize;
        else if (kind == 'd') ok = (*f == 'd');
        else if (kind == 'f') ok = (*f == 'f');
    }
    if (!ok) {
        PyBuffer_Release(&view);
        PyErr_SetString(PyExc_TypeError, "array has the wrong dtype for the native ingest path");
        return -1;
    }
    *addr = (uintptr_t)view.buf;
    *count = view.itemsize ? view.len / view.itemsize : 0;
    PyBuffer_Release(&view); /* the array object stays alive through `keepalive` */
    return 0;
}

/* item.name through the generic attribute protocol.  (Reading the instance __dict__ directly is no faster on
 * CPython 3.12: objects keep their attributes inline until somebody asks for the dict.)  New reference. */
static PyObject *get_field(PyObject *item, PyObject *name) { return PyObject_GetAttr(item, name); }

static PyObject *collect(PyObject *self, PyObject *args) {
    PyObject *seq_in;
    int want_data = 0;
    if (!PyArg_ParseTuple(args, "O|p", &seq_in, &want_data)) return NULL;
    PyObject *seq = PySequence_Fast(seq_in, "expected a sequence of sparse matrices");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject *t_ip = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(uintptr_t));
    PyObject *t_ix = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(uintptr_t));
    PyObject *t_d = want_data ? PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(uintptr_t)) : NULL;
    PyObject *t_n = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(int32_t));
    PyObject *keep = PyList_New(0);
    if (!t_ip || !t_ix || (want_data && !t_d) || !t_n || !keep) goto fail;
    {
        uintptr_t *ip = (uintptr_t *)PyBytes_AS_STRING(t_ip), *ix = (uintptr_t *)PyBytes_AS_STRING(t_ix);
        uintptr_t *dp = want_data ? (uintptr_t *)PyBytes_AS_STRING(t_d) : NULL;
        int32_t *nr = (int32_t *)PyBytes_AS_STRING(t_n);
        long long total = 0;
        for (Py_ssize_t g = 0; g < n; ++g) {
            PyObject *item = PySequence_Fast_GET_ITEM(seq, g); /* borrowed */
            PyObject *a = get_field(item, s_indptr);
            if (!a) goto type_fail;
            PyObject *b = get_field(item, s_indices);
            if (!b) {
                Py_DECREF(a);
                goto type_fail;
            }
            Py_ssize_t ca = 0, cb = 0;
            int bad = buffer_of(a, 4, 'i', &ip[g], &ca) != 0 || buffer_of(b, 4, 'i', &ix[g], &cb) != 0 || ca < 1;
            if (!bad) bad = PyList_Append(keep, a) != 0 || PyList_Append(keep, b) != 0;
            Py_DECREF(a);
            Py_DECREF(b);
            if (bad) goto type_fail;
            nr[g] = (int32_t)(ca - 1);
            total += ca - 1;
            if (want_data) {
                PyObject *d = get_field(item, s_data);
                Py_ssize_t cd = 0;
                if (!d) goto type_fail;
                bad = buffer_of(d, 8, 'd', &dp[g], &cd) != 0 || cd < cb || PyList_Append(keep, d) != 0;
                Py_DECREF(d);
                if (bad) goto type_fail;
            }
        }
        Py_DECREF(seq);
        if (!want_data) {
            t_d = Py_None;
            Py_INCREF(Py_None);
        }
        return Py_BuildValue("(NNNNNL)", t_ip, t_ix, t_d, t_n, keep, total);
    }
type_fail:
    if (!PyErr_Occurred() || PyErr_ExceptionMatches(PyExc_AttributeError) || PyErr_ExceptionMatches(PyExc_BufferError)) {
        PyErr_Clear();
        PyErr_SetString(PyExc_TypeError, "item is not a CSR/CSC matrix with int32 indptr / indices (and float64 data)");
    }
fail:
    Py_XDECREF(seq);
    Py_XDECREF(t_ip);
    Py_XDECREF(t_ix);
    Py_XDECREF(t_d);
    Py_XDECREF(t_n);
    Py_XDECREF(keep);
    return NULL;
}

static PyObject *pointers(PyObject *self, PyObject *args) {
    PyObject *seq_in;
    int itemsize = 8;
    if (!PyArg_ParseTuple(args, "O|i", &seq_in, &itemsize)) return NULL;
    PyObject *seq = PySequence_Fast(seq_in, "expected a sequence of arrays");
    if (!seq) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(seq);
    PyObject *tab = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(uintptr_t));
    PyObject *len = PyBytes_FromStringAndSize(NULL, n * (Py_ssize_t)sizeof(int32_t));
    PyObject *keep = PyList_New(0);
    if (!tab || !len || !keep) goto fail;
    {
        uintptr_t *p = (uintptr_t *)PyBytes_AS_STRING(tab);
        int32_t *l = (int32_t *)PyBytes_AS_STRING(len);
        for (Py_ssize_t g = 0; g < n; ++g) {
            PyObject *item = PySequence_Fast_GET_ITEM(seq, g);
            Py_ssize_t c = 0;
            if (buffer_of(item, itemsize, itemsize == 8 ? 'd' : 'x', &p[g], &c) != 0) goto fail;
            if (PyList_Append(keep, item) != 0) goto fail;
            l[g] = (int32_t)c;
        }
        Py_DECREF(seq);
        return Py_BuildValue("(NNN)", tab, len, keep);
    }
fail:
    Py_XDECREF(seq);
    Py_XDECREF(tab);
    Py_XDECREF(len);
    Py_XDECREF(keep);
    return NULL;
}

static PyMethodDef methods[] = {
    {"collect", collect, METH_VARARGS, "collect(seq, want_data=False) -> pointer tables of a list of CSR/CSC matrices"},
    {"pointers", pointers, METH_VARARGS, "pointers(seq, itemsize=8) -> pointer table of a list of contiguous arrays"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_pyingest", "pointer tables for the native ingest path", -1, methods};

PyMODINIT_FUNC PyInit__pyingest(void) {
    s_indptr = PyUnicode_InternFromString("indptr");
    s_indices = PyUnicode_InternFromString("indices");
    s_data = PyUnicode_InternFromString("data");
    if (!s_indptr || !s_indices || !s_data) return NULL;
    return PyModule_Create(&moduledef);
}
