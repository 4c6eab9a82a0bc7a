/* hoststage.c -- CPython helper for ReplayBuffer.add (General/Base/replay_buffer.py:58-65 in the reference).
 *
 * The reference's add() is five numpy __setitem__ calls per transition (~1 us); the drop-in stages transitions in host
 * arrays and hands them to the library at the next train step, so add() is pure host bookkeeping on the e2e path.  This
 * module does the five stores in C through the buffer protocol (~0.3 us).  It is an optional accelerator of HOST staging
 * only: without it replay.py uses the numpy stores; no device work lives here.
 *
 *   stage = hoststage.new(addr_states, addr_actions, addr_rewards, addr_observations, addr_dones, D, capacity)
 *   hoststage.put(stage, i, state, action, reward, observation, done)     # row i of the five staging arrays
 *
 * It can also make the two C-ABI calls of the reference's per-step sequence without going through ctypes (~1.2 us of
 * argument marshalling each): after hoststage.bind(stage, handle, &dqn_store_train_step, &dqn_get_losses, &dqn_get_loss_lagged, agent),
 *   hoststage.step(stage, n)     == dqn_store_train_step(handle, agent, n, <the staging arrays>, 1, NULL)   -> status
 *   hoststage.loss(stage)        == dqn_get_losses(handle, agent, 1, &loss, NULL)                             -> (status, loss)
 *   hoststage.loss(stage, 1)     == dqn_get_loss_lagged(handle, agent, 1, &loss)                              -> (status, loss)
 * The functions are the library's own entry points (include/dqn_b200.h), called by address.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  float* s;
  int64_t* a;
  float* r;
  float* o;
  uint8_t* d;
  Py_ssize_t D, cap;
  /* optional direct binding of the library (hs_bind) */
  void* handle;
  int (*store_train_step)(void*, int32_t, int64_t, const float*, const int64_t*, const float*, const float*, const uint8_t*, int32_t, float*);
  int (*get_losses)(void*, int32_t, int32_t, float*, int64_t*);
  int (*get_loss_lagged)(void*, int32_t, int32_t, float*);
  int32_t agent;
} Stage;

static void stage_free(PyObject* cap) { PyMem_Free(PyCapsule_GetPointer(cap, "dqn_b200.hoststage")); }

static PyObject* hs_new(PyObject* self, PyObject* args) {
  unsigned long long as, aa, ar, ao, ad;
  Py_ssize_t D, cap;
  if (!PyArg_ParseTuple(args, "KKKKKnn", &as, &aa, &ar, &ao, &ad, &D, &cap)) return NULL;
  Stage* st = (Stage*)PyMem_Malloc(sizeof(Stage));
  if (!st) return PyErr_NoMemory();
  st->s = (float*)(uintptr_t)as; st->a = (int64_t*)(uintptr_t)aa; st->r = (float*)(uintptr_t)ar;
  st->o = (float*)(uintptr_t)ao; st->d = (uint8_t*)(uintptr_t)ad; st->D = D; st->cap = cap;
  st->handle = NULL; st->store_train_step = NULL; st->get_losses = NULL; st->get_loss_lagged = NULL; st->agent = 0;
  return PyCapsule_New(st, "dqn_b200.hoststage", stage_free);
}

/* copy D float32 values out of any C-contiguous float32 buffer (numpy row, memoryview, ...) */
static int copy_obs(PyObject* obj, float* dst, Py_ssize_t D) {
  Py_buffer view;
  if (PyObject_GetBuffer(obj, &view, PyBUF_C_CONTIGUOUS | PyBUF_FORMAT) != 0) return -1;
  const int ok = view.len == D * 4 && view.format && view.format[0] == 'f' && view.format[1] == 0;
  if (ok) memcpy(dst, view.buf, (size_t)D * 4);
  PyBuffer_Release(&view);
  if (!ok) {
    PyErr_SetString(PyExc_TypeError, "hoststage.put: observation must be a contiguous float32 buffer of D values");
    return -1;
  }
  return 0;
}

static PyObject* hs_put(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  if (nargs != 7) { PyErr_SetString(PyExc_TypeError, "put(stage, i, state, action, reward, observation, done)"); return NULL; }
  Stage* st = (Stage*)PyCapsule_GetPointer(args[0], "dqn_b200.hoststage");
  if (!st) return NULL;
  const Py_ssize_t i = PyLong_AsSsize_t(args[1]);
  if (i < 0 || i >= st->cap) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "hoststage.put: row out of range"); return NULL; }
  const long long action = PyLong_AsLongLong(args[3]);           /* numpy integer scalars go through __index__ */
  if (action == -1 && PyErr_Occurred()) return NULL;
  const double reward = PyFloat_AsDouble(args[4]);
  if (reward == -1.0 && PyErr_Occurred()) return NULL;
  const int done = PyObject_IsTrue(args[6]);
  if (done < 0) return NULL;
  if (copy_obs(args[2], st->s + i * st->D, st->D) != 0) return NULL;
  if (copy_obs(args[5], st->o + i * st->D, st->D) != 0) return NULL;
  st->a[i] = (int64_t)action;
  st->r[i] = (float)reward;                                      /* the reference stores into a float32 array */
  st->d[i] = (uint8_t)done;
  Py_RETURN_NONE;
}

static PyObject* hs_bind(PyObject* self, PyObject* args) {
  PyObject* cap;
  unsigned long long handle, f_step, f_loss, f_lagged;
  int agent;
  if (!PyArg_ParseTuple(args, "OKKKKi", &cap, &handle, &f_step, &f_loss, &f_lagged, &agent)) return NULL;
  Stage* st = (Stage*)PyCapsule_GetPointer(cap, "dqn_b200.hoststage");
  if (!st) return NULL;
  st->handle = (void*)(uintptr_t)handle;
  *(void**)(&st->store_train_step) = (void*)(uintptr_t)f_step;
  *(void**)(&st->get_losses) = (void*)(uintptr_t)f_loss;
  *(void**)(&st->get_loss_lagged) = (void*)(uintptr_t)f_lagged;
  st->agent = (int32_t)agent;
  Py_RETURN_NONE;
}

static PyObject* hs_step(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  if (nargs != 2) { PyErr_SetString(PyExc_TypeError, "step(stage, n)"); return NULL; }
  Stage* st = (Stage*)PyCapsule_GetPointer(args[0], "dqn_b200.hoststage");
  if (!st) return NULL;
  if (!st->store_train_step) { PyErr_SetString(PyExc_RuntimeError, "hoststage.step: bind() has not been called"); return NULL; }
  const Py_ssize_t n = PyLong_AsSsize_t(args[1]);
  if (n < 0 || n > st->cap) { if (!PyErr_Occurred()) PyErr_SetString(PyExc_IndexError, "hoststage.step: n out of range"); return NULL; }
  return PyLong_FromLong(st->store_train_step(st->handle, st->agent, (int64_t)n, st->s, st->a, st->r, st->o, st->d, 1, NULL));
}

static PyObject* hs_loss(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  if (nargs != 1 && nargs != 2) { PyErr_SetString(PyExc_TypeError, "loss(stage[, lag])"); return NULL; }
  Stage* st = (Stage*)PyCapsule_GetPointer(args[0], "dqn_b200.hoststage");
  if (!st) return NULL;
  if (!st->get_losses) { PyErr_SetString(PyExc_RuntimeError, "hoststage.loss: bind() has not been called"); return NULL; }
  long lag = 0;
  if (nargs == 2) { lag = PyLong_AsLong(args[1]); if (lag == -1 && PyErr_Occurred()) return NULL; }
  float loss = 0.f;
  const int rc = lag ? st->get_loss_lagged(st->handle, st->agent, (int32_t)lag, &loss) : st->get_losses(st->handle, st->agent, 1, &loss, NULL);
  return Py_BuildValue("(id)", rc, (double)loss);
}

static PyMethodDef methods[] = {
    {"new", hs_new, METH_VARARGS, "new(addr_s, addr_a, addr_r, addr_o, addr_d, D, capacity) -> stage"},
    {"put", (PyCFunction)(void (*)(void))hs_put, METH_FASTCALL, "put(stage, i, state, action, reward, observation, done)"},
    {"bind", hs_bind, METH_VARARGS, "bind(stage, handle, addr_dqn_store_train_step, addr_dqn_get_losses, addr_dqn_get_loss_lagged, agent)"},
    {"step", (PyCFunction)(void (*)(void))hs_step, METH_FASTCALL, "step(stage, n) -> status of dqn_store_train_step(handle, agent, n, staged arrays, 1, NULL)"},
    {"loss", (PyCFunction)(void (*)(void))hs_loss, METH_FASTCALL, "loss(stage[, lag]) -> (status, loss of the last train step / of the one before it)"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hoststage", "host-side staging of ReplayBuffer.add", -1, methods};

PyMODINIT_FUNC PyInit__hoststage(void) { return PyModule_Create(&moddef); }
