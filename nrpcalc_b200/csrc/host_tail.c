/* Host tail of the Finder graph build as a CPython extension: the SAME container operations as
 * nrpcalc_b200/hgraph.py (replay_cliques, _maximal_cliques, build_homology_graph), i.e. as the reference's
 * nrpcalc/base/hgraph.py:96-200, issued through the C API on REAL CPython set / dict / tuple / deque objects.
 *
 * The order of the resulting graph depends on CPython's set layout (tuple hashes, probing, resizes, the pop() finger), so
 * nothing is re-implemented here: every `set.add`, `set.pop`, `|=`, `&=`, `-`, `update`, `difference`, `tuple(set)` below is
 * the interpreter's own routine and the result is identical by construction, on any CPython version.  What goes away is the
 * bytecode loop around them (2-3x on this stage).  hgraph.py falls back to its Python loops when this module is not built.
 *
 * Plain C, no CUDA: built by `make -C nrpcalc_b200/csrc host` (gcc) into nrpcalc_b200/_host_tail.so.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

static PyObject *deque_type;       /* collections.deque */
static PyObject *str_update, *str_difference, *str_popleft;

/* hgraph.py:96-115 -- cliques that are not contained in another one.  `cliques` is a list of sets. */
static PyObject *maximal_cliques(PyObject *cliques) {
  const Py_ssize_t n = PyList_GET_SIZE(cliques);
  PyObject *member_of = PyDict_New(), *keep = NULL, *result = NULL;
  if (!member_of) return NULL;
  /* member_of[node].add(pos) for every node of every clique, cliques in list order, nodes in set order */
  for (Py_ssize_t pos = 0; pos < n; ++pos) {
    PyObject *pos_obj = PyLong_FromSsize_t(pos);
    PyObject *it = pos_obj ? PyObject_GetIter(PyList_GET_ITEM(cliques, pos)) : NULL;
    PyObject *node;
    if (!it) { Py_XDECREF(pos_obj); goto fail; }
    while ((node = PyIter_Next(it)) != NULL) {
      PyObject *where = PyDict_GetItemWithError(member_of, node);      /* borrowed */
      int rc = 0;
      if (!where) {
        if (PyErr_Occurred()) rc = -1;
        else {
          where = PySet_New(NULL);
          if (!where || PyDict_SetItem(member_of, node, where) < 0) rc = -1;
          Py_XDECREF(where);                                           /* the dict keeps it alive */
        }
      }
      if (rc == 0) rc = PySet_Add(where, pos_obj);
      Py_DECREF(node);
      if (rc < 0) { Py_DECREF(it); Py_DECREF(pos_obj); goto fail; }
    }
    Py_DECREF(it);
    Py_DECREF(pos_obj);
    if (PyErr_Occurred()) goto fail;
  }
  keep = PySet_New(NULL);
  if (!keep) goto fail;
  for (Py_ssize_t pos = 0; pos < n; ++pos) {
    PyObject *pos_obj = PyLong_FromSsize_t(pos);
    PyObject *common = PySet_New(NULL);
    PyObject *it = (pos_obj && common) ? PyObject_GetIter(PyList_GET_ITEM(cliques, pos)) : NULL;
    PyObject *node;
    int rc = it ? 0 : -1;
    while (rc == 0 && (node = PyIter_Next(it)) != NULL) {
      PyObject *where = PyDict_GetItemWithError(member_of, node);      /* borrowed; always present */
      PyObject *updated = NULL;
      if (where) updated = PySet_GET_SIZE(common) == 0 ? PyNumber_InPlaceOr(common, where)     /* common |= member_of[node] */
                                                        : PyNumber_InPlaceAnd(common, where);   /* common &= member_of[node] */
      Py_DECREF(node);
      if (!updated) { rc = -1; break; }
      Py_SETREF(common, updated);
    }
    Py_XDECREF(it);
    if (rc == 0 && PyErr_Occurred()) rc = -1;
    if (rc == 0) {
      if (PySet_GET_SIZE(common) == 1) rc = PySet_Add(keep, pos_obj);
      else {                                                           /* keep.update(common - set([pos])) */
        PyObject *self = PySet_New(NULL), *rest = NULL, *none = NULL;
        if (self && PySet_Add(self, pos_obj) == 0) rest = PyNumber_Subtract(common, self);
        if (rest) none = PyObject_CallMethodOneArg(keep, str_update, rest);
        if (!none) rc = -1;
        Py_XDECREF(self); Py_XDECREF(rest); Py_XDECREF(none);
      }
    }
    Py_XDECREF(common);
    Py_XDECREF(pos_obj);
    if (rc < 0) goto fail;
  }
  /* [cliques[pos] for pos in keep] */
  result = PyList_New(0);
  if (!result) goto fail;
  {
    PyObject *it = PyObject_GetIter(keep), *pos_obj;
    if (!it) goto fail;
    while ((pos_obj = PyIter_Next(it)) != NULL) {
      const Py_ssize_t pos = PyLong_AsSsize_t(pos_obj);
      Py_DECREF(pos_obj);
      if (PyList_Append(result, PyList_GET_ITEM(cliques, pos)) < 0) { Py_DECREF(it); goto fail; }
    }
    Py_DECREF(it);
    if (PyErr_Occurred()) goto fail;
  }
  Py_DECREF(member_of);
  Py_DECREF(keep);
  return result;
fail:
  Py_XDECREF(member_of); Py_XDECREF(keep); Py_XDECREF(result);
  return NULL;
}

/* replay_cliques(tuples) -> list of deques (hgraph.py:130-155 on the compressed stream of distinct id tuples) */
static PyObject *replay_cliques(PyObject *module, PyObject *tuples) {
  PyObject *distinct = PySet_New(NULL), *as_sets = NULL, *kept = NULL, *final = NULL, *out = NULL, *it = NULL, *item;
  if (!distinct) return NULL;
  it = PyObject_GetIter(tuples);
  if (!it) goto fail;
  while ((item = PyIter_Next(it)) != NULL) {                           /* distinct.add(members) */
    const int rc = PySet_Add(distinct, item);
    Py_DECREF(item);
    if (rc < 0) goto fail;
  }
  Py_CLEAR(it);
  if (PyErr_Occurred()) goto fail;
  as_sets = PyList_New(0);
  if (!as_sets) goto fail;
  while (PySet_GET_SIZE(distinct) > 0) {                               /* as_sets.append(set(distinct.pop())) */
    PyObject *members = PySet_Pop(distinct);
    PyObject *as_set = members ? PySet_New(members) : NULL;
    const int rc = as_set ? PyList_Append(as_sets, as_set) : -1;
    Py_XDECREF(members); Py_XDECREF(as_set);
    if (rc < 0) goto fail;
  }
  kept = maximal_cliques(as_sets);
  if (!kept) goto fail;
  Py_CLEAR(as_sets);
  final = PySet_New(NULL);
  if (!final) goto fail;
  for (Py_ssize_t i = PyList_GET_SIZE(kept); i-- > 0;) {               /* final.add(tuple(as_sets.pop())) */
    PyObject *as_tuple = PySequence_Tuple(PyList_GET_ITEM(kept, i));
    const int rc = as_tuple ? PySet_Add(final, as_tuple) : -1;
    Py_XDECREF(as_tuple);
    if (rc < 0) goto fail;
    if (PyList_SetSlice(kept, i, i + 1, NULL) < 0) goto fail;          /* the pop: the clique's set dies here, as in Python */
  }
  out = PyList_New(0);
  if (!out) goto fail;
  while (PySet_GET_SIZE(final) > 0) {                                  /* yield deque(final.pop()) */
    PyObject *members = PySet_Pop(final);
    PyObject *clique = members ? PyObject_CallOneArg(deque_type, members) : NULL;
    const int rc = clique ? PyList_Append(out, clique) : -1;
    Py_XDECREF(members); Py_XDECREF(clique);
    if (rc < 0) goto fail;
  }
  Py_DECREF(distinct); Py_DECREF(kept); Py_DECREF(final);
  return out;
fail:
  Py_XDECREF(distinct); Py_XDECREF(as_sets); Py_XDECREF(kept); Py_XDECREF(final); Py_XDECREF(out); Py_XDECREF(it);
  return NULL;
}

/* node in the graph's dicts, as nx.Graph.add_node / add_edge create it for an attribute-free graph */
static int ensure_node(PyObject *adj, PyObject *nodes, PyObject *u) {
  const int has = PyDict_Contains(nodes, u);
  if (has != 0) return has < 0 ? -1 : 0;
  PyObject *inner = PyDict_New(), *attr = PyDict_New();
  int rc = (inner && attr) ? 0 : -1;
  if (rc == 0) rc = PyDict_SetItem(adj, u, inner);
  if (rc == 0) rc = PyDict_SetItem(nodes, u, attr);
  Py_XDECREF(inner); Py_XDECREF(attr);
  return rc;
}

/* fill_graph(adj, nodes, cliques) -> number of cliques inserted (hgraph.py:157-200; adj / nodes are graph._adj / graph._node) */
static PyObject *fill_graph(PyObject *module, PyObject *args) {
  PyObject *adj, *nodes, *cliques, *it, *clique;
  Py_ssize_t count = 0;
  if (!PyArg_ParseTuple(args, "O!O!O", &PyDict_Type, &adj, &PyDict_Type, &nodes, &cliques)) return NULL;
  it = PyObject_GetIter(cliques);
  if (!it) return NULL;
  while ((clique = PyIter_Next(it)) != NULL) {
    int rc = 0;
    Py_ssize_t len = PyObject_Length(clique);
    if (len < 0) rc = -1;
    else if (len == 1) {                                               /* graph.add_node(clique.popleft()) */
      PyObject *u = PyObject_CallMethodNoArgs(clique, str_popleft);
      rc = u ? ensure_node(adj, nodes, u) : -1;
      Py_XDECREF(u);
    } else {
      PyObject *pending = PySet_New(clique);                           /* pending = set(clique) */
      if (!pending) rc = -1;
      while (rc == 0 && (len = PyObject_Length(clique)) > 0) {
        PyObject *u = PyObject_CallMethodNoArgs(clique, str_popleft);
        PyObject *nbrs = NULL, *fresh = NULL, *vit = NULL, *v;
        if (!u) { rc = -1; break; }
        if (PySet_Discard(pending, u) < 0) rc = -1;                    /* pending.remove(u): u is always a member */
        if (rc == 0) rc = ensure_node(adj, nodes, u);
        if (rc == 0) nbrs = PyDict_GetItemWithError(adj, u);           /* borrowed */
        if (rc == 0 && !nbrs) rc = -1;
        if (rc == 0) fresh = PyObject_CallMethodOneArg(pending, str_difference, nbrs);   /* pending - set(graph[u]) */
        if (rc == 0 && !fresh) rc = -1;
        if (rc == 0) vit = PyObject_GetIter(fresh);
        if (rc == 0 && !vit) rc = -1;
        while (rc == 0 && (v = PyIter_Next(vit)) != NULL) {            /* graph.add_edge(u, v) */
          PyObject *data = NULL, *back;
          rc = ensure_node(adj, nodes, v);
          if (rc == 0) data = PyDict_New();
          if (rc == 0 && !data) rc = -1;
          if (rc == 0) rc = PyDict_SetItem(nbrs, v, data);
          if (rc == 0) { back = PyDict_GetItemWithError(adj, v); rc = back ? PyDict_SetItem(back, u, data) : -1; }
          Py_XDECREF(data);
          Py_DECREF(v);
        }
        if (rc == 0 && PyErr_Occurred()) rc = -1;
        Py_XDECREF(vit); Py_XDECREF(fresh); Py_DECREF(u);
      }
      if (len < 0) rc = -1;
      Py_XDECREF(pending);
    }
    Py_DECREF(clique);
    if (rc < 0) { Py_DECREF(it); return NULL; }
    ++count;
  }
  Py_DECREF(it);
  if (PyErr_Occurred()) return NULL;
  return PyLong_FromSsize_t(count);
}

static PyMethodDef methods[] = {
    {"replay_cliques", replay_cliques, METH_O, "replay_cliques(tuples) -> list of deques (hgraph.py:130-155)"},
    {"fill_graph", fill_graph, METH_VARARGS, "fill_graph(adj, nodes, cliques) -> cliques inserted (hgraph.py:157-200)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moduledef = {PyModuleDef_HEAD_INIT, "_host_tail", "Finder host tail on real CPython containers", -1, methods};

PyMODINIT_FUNC PyInit__host_tail(void) {
  PyObject *collections = PyImport_ImportModule("collections");
  if (!collections) return NULL;
  deque_type = PyObject_GetAttrString(collections, "deque");
  Py_DECREF(collections);
  str_update = PyUnicode_InternFromString("update");
  str_difference = PyUnicode_InternFromString("difference");
  str_popleft = PyUnicode_InternFromString("popleft");
  if (!deque_type || !str_update || !str_difference || !str_popleft) return NULL;
  return PyModule_Create(&moduledef);
}
