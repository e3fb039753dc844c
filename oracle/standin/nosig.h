// Force-included (g++ -include) when building oracle/_ref: the reference calls
// PyErr_CheckSignals() from OpenMP worker threads (src/pairsnp.hpp:385), which dereferences a
// NULL thread state on CPython 3.12 and segfaults for n_threads > 1. Reference sources stay unmodified.
#include <Python.h>
#undef PyErr_CheckSignals
#define PyErr_CheckSignals() 0
