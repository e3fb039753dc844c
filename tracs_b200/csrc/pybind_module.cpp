// Compiled extension module `TRACS` over the C ABI of libtracs_b200.so: the pybind11 face of the drop-in
// (option B). Same module name, function names, keyword names and return types as the reference's
// module (gtonkinhill/tracs src/python_bindings.cpp:8-26), so `from TRACS import pairsnp` in tracs/distance.py:8,
// `trans_dist` in tracs/transcluster.py:2, `calculate_posteriors` in tracs/align.py:21 and `lprob_k_given_N` in
// tests/test_llk.py:3 bind to it unchanged. No arithmetic happens here: every call goes to include/tracs_b200.h.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <stdio.h>

#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/tracs_b200.h"

namespace py = pybind11;

namespace {

void check(int rc) {
  if (rc == 0) return;
  const std::string msg = tracs_last_error();
  if (rc == 3) throw py::index_error(msg);  // std::out_of_range -> IndexError, like pybind11's own translation
  if (rc == 4) {                            // SIGINT during the call: the reference prints this and exits with status 1
    fprintf(stderr, "Interrupted by user!\n");
    PyErr_SetObject(PyExc_SystemExit, py::int_(1).ptr());
    throw py::error_already_set();
  }
  throw std::runtime_error(msg);            // -> RuntimeError
}

template <typename T>
py::list to_list(const T *p, size_t n) {
  py::list out(n);
  for (size_t i = 0; i < n; ++i) out[i] = p ? p[i] : T(0);
  return out;
}

// src/pairsnp.hpp:320-322, 457: (rows, cols, distances, seq_names, filt_distances, n_compared_sites), all lists
py::tuple pairsnp(const std::vector<std::string> &fasta, int n_threads, int dist, bool filter) {
  std::vector<const char *> paths;
  for (const std::string &f : fasta) paths.push_back(f.c_str());
  if (paths.empty()) paths.push_back("");
  tracs_edges_t e;
  int rc;
  {
    py::gil_scoped_release release;  // nothing below touches Python objects
    rc = tracs_pairsnp(paths.data(), (int)fasta.size(), n_threads, (int32_t)dist, filter ? 1 : 0, &e);
  }
  check(rc);
  py::list names(e.n_names);
  for (size_t i = 0; i < e.n_names; ++i) names[i] = py::str(e.names[i]);
  py::tuple out = py::make_tuple(to_list(e.rows, e.n_edges), to_list(e.cols, e.n_edges), to_list(e.dist, e.n_edges), names,
                                 to_list(e.filt, e.n_edges), to_list(e.ncomp, e.n_edges));
  tracs_edges_free(&e);
  return out;
}

// src/transcluster.hpp:240-287: (log P(direct transmission) per edge, E[K] per edge)
py::tuple trans_dist(const std::vector<long long> &snpdiff, const std::vector<double> &datediff, double lamb, double beta,
                     double threshold_Ek) {
  if (snpdiff.size() != datediff.size()) throw py::index_error("snpdiff and datediff must have the same length");
  std::vector<int32_t> snp(snpdiff.begin(), snpdiff.end());
  std::vector<double> p0(snp.size()), eK(snp.size());
  check(tracs_trans_dist(snp.data(), datediff.data(), snp.size(), lamb, beta, threshold_Ek, p0.data(), eK.data()));
  return py::make_tuple(to_list(p0.data(), p0.size()), to_list(eK.data(), eK.size()));
}

// src/transcluster.hpp:90-129
py::tuple lprob_k_given_N(size_t N, size_t k, double delta, double lamb, double beta, const std::vector<double> &lgamma) {
  double out[2];
  check(tracs_lprob_k_given_N(N, k, delta, lamb, beta, lgamma.data(), lgamma.size(), out));
  return py::make_tuple(out[0], out[1]);
}

// src/dmultinomial.hpp:8-86
py::array_t<double> calculate_posteriors(py::array_t<double, py::array::c_style | py::array::forcecast> counts,
                                         const std::vector<double> &alphas, bool keep, double threshold) {
  if (counts.ndim() != 2) throw std::runtime_error("counts must be a 2-D array");
  py::array_t<double> out({counts.shape(0), counts.shape(1)});
  check(tracs_calculate_posteriors(counts.data(), (size_t)counts.shape(0), (size_t)counts.shape(1), alphas.data(), alphas.size(),
                                   keep ? 1 : 0, threshold, out.mutable_data()));
  return out;
}

}  // namespace

PYBIND11_MODULE(TRACS, m) {
  m.doc() = "Meta Transmission Clustering";
  m.def("pairsnp", &pairsnp, py::arg("fasta"), py::arg("n_threads"), py::arg("dist"), py::arg("filter"));
  m.def("lprob_k_given_N", &lprob_k_given_N, py::arg("N"), py::arg("k"), py::arg("delta"), py::arg("lamb"), py::arg("beta"),
        py::arg("lgamma"));
  m.def("trans_dist", &trans_dist, py::arg("snpdiff"), py::arg("datediff"), py::arg("lamb"), py::arg("beta"), py::arg("threshold_Ek"));
  m.def("calculate_posteriors", &calculate_posteriors, py::arg("counts"), py::arg("alphas"), py::arg("keep"), py::arg("threshold"));
}
