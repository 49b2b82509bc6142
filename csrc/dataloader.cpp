// Native batch assembler: multi-threaded random-crop gather from a uint16 token memmap into int32
// (pinned) host buffers.  Replaces the reference's numpy fancy-index `np.take` (src/train.py:61-62),
// which is single-threaded and allocates two temporaries per batch.
#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>

#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace py = pybind11;

static void gather_crops(py::array data, py::array_t<int64_t, py::array::c_style | py::array::forcecast> starts, int64_t T,
                         py::array_t<int32_t, py::array::c_style> x_out, py::array_t<int32_t, py::array::c_style> y_out,
                         int n_threads) {
  if (data.itemsize() != 2 || data.ndim() != 1) throw std::runtime_error("data must be a 1-D uint16 array");
  const uint16_t* src = static_cast<const uint16_t*>(data.data());
  const int64_t len = data.shape(0);
  const int64_t n = starts.shape(0);
  if (x_out.ndim() != 2 || x_out.shape(0) != n || x_out.shape(1) != T || y_out.shape(0) != n || y_out.shape(1) != T)
    throw std::runtime_error("x_out / y_out must be [n, T] int32");
  const int64_t* st = starts.data();
  for (int64_t i = 0; i < n; ++i)
    if (st[i] < 0 || st[i] + T + 1 > len) throw std::runtime_error("crop out of range");
  int32_t* xo = x_out.mutable_data();
  int32_t* yo = y_out.mutable_data();
  py::gil_scoped_release release;
  const int nt = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(n_threads, n)));
  auto work = [&](int tid) {
    for (int64_t i = tid; i < n; i += nt) {
      const uint16_t* s = src + st[i];
      int32_t* xr = xo + i * T;
      int32_t* yr = yo + i * T;
      for (int64_t t = 0; t < T; ++t) xr[t] = s[t];          // two independent widening copies: both auto-vectorise
      for (int64_t t = 0; t < T; ++t) yr[t] = s[t + 1];
    }
  };
  if (nt == 1) { work(0); return; }
  std::vector<std::thread> pool;
  for (int t = 0; t < nt; ++t) pool.emplace_back(work, t);
  for (auto& th : pool) th.join();
}

PYBIND11_MODULE(_native_loader, m) {
  m.def("gather_crops", &gather_crops, py::arg("data"), py::arg("starts"), py::arg("T"), py::arg("x_out"), py::arg("y_out"),
        py::arg("n_threads") = 8);
}
