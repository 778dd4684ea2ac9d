// python_bindings.cpp — the `window_ann` pybind11 module of this engine.
//
// Mirrors the part of the reference module that experiments/wrapper.py and
// experiments/run_our_method.py use (python_bindings/python_bindings.cpp:111-157,204-213):
// the same class names, constructor keyword arguments, batch_search argument order and
// return type (tuple of uint32 ids [nq,k] and float32 distances [nq,k]).  The float variants are
// the ones wrapper.py reaches; the reference also registers UInt8 / Int8 variants
// (python_bindings.cpp:74-86,233-236), provided here over the same fp32 arena: their distances are
// integer sums (euclidian_point.h:44-60, mips_point.h:44-58) cast to float, which fp32 arithmetic
// reproduces exactly, in any summation order, while every partial sum stays below 2^24 — checked at
// construction (dim * 255^2, or dim * 128^2 for int8 MIPS).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "window_index.hpp"

namespace py = pybind11;
using namespace pybind11::literals;
using wsann::BuildParams;
using wsann::QueryParams;

using FArray = py::array_t<float, py::array::c_style | py::array::forcecast>;
template <class T>
using TArray = py::array_t<T, py::array::c_style | py::array::forcecast>;
using Result = std::pair<py::array_t<unsigned int>, py::array_t<float>>;

static wsann::BuildParams DEFAULT_BUILD_PARAMS(64, 500, 1.175, "index_cache");  // python_bindings.cpp:88

// fp32 view of a [rows][dim] array of T: the array itself for float, a widened copy otherwise
template <class T>
struct Widened {
  std::vector<float> copy;
  const float* ptr = nullptr;
  Widened(const T* src, size_t count) {
    if constexpr (std::is_same<T, float>::value) {
      ptr = src;
    } else {
      copy.resize(count);
      for (size_t i = 0; i < count; i++) copy[i] = (float)src[i];
      ptr = copy.data();
    }
  }
};

template <class T>
struct Points {
  Widened<T> w;
  const float* data;
  size_t n, dim;
};

// integer variants: the exactness domain of the fp32 arena (see the header comment)
template <class T, int METRIC>
static void check_exact_domain(size_t dim) {
  if constexpr (!std::is_same<T, float>::value) {
    const uint64_t term = (std::is_same<T, int8_t>::value && METRIC == WS_METRIC_MIPS) ? 128ull * 128ull : 255ull * 255ull;
    if ((uint64_t)dim * term >= (1ull << 24))
      throw std::runtime_error("8-bit variants are supported up to " + std::to_string(((1ull << 24) - 1) / term) +
                               " dimensions (integer distances must stay exactly representable in fp32)");
  }
}

// prefiltering.h:78-98 / tree_utils.h:44-60 error behaviour
template <class T, int METRIC>
static Points<T> check_points(const TArray<T>& points, const FArray& filter_values) {
  if (points.ndim() != 2) throw std::runtime_error("points numpy array must be 2-dimensional");
  if (filter_values.ndim() != 1) throw std::runtime_error("filter data numpy array must be 1-dimensional");
  if (filter_values.shape(0) != points.shape(0))
    throw std::runtime_error("filter data numpy array must have the same number of elements as the points array");
  check_exact_domain<T, METRIC>((size_t)points.shape(1));
  Points<T> p{Widened<T>(points.data(), (size_t)points.size()), nullptr, (size_t)points.shape(0), (size_t)points.shape(1)};
  p.data = p.w.ptr;
  return p;
}

template <class T>
struct Batch {
  Widened<T> w;
  const float* queries;
  const float* filters;
  uint64_t nq;
};

template <class T>
static Batch<T> check_batch(const TArray<T>& queries, const FArray& filters, uint64_t num_queries, size_t dim) {
  if (queries.ndim() != 2 || (size_t)queries.shape(1) != dim)
    throw std::runtime_error("queries must be a 2-dimensional array with the index's dimension");
  if ((uint64_t)queries.shape(0) < num_queries) throw std::runtime_error("fewer query rows than num_queries");
  if (filters.ndim() != 2 || filters.shape(1) != 2) throw std::runtime_error("filters must be a sequence of (lo, hi) pairs");
  if ((uint64_t)filters.shape(0) < num_queries) throw std::runtime_error("fewer filters than num_queries");
  Batch<T> b{Widened<T>(queries.data(), (size_t)num_queries * dim), nullptr, filters.data(), num_queries};
  b.queries = b.w.ptr;
  return b;
}

template <class F>
static Result run(uint64_t nq, long k, F&& f) {
  if (k < 1) throw std::runtime_error("k must be >= 1");
  py::array_t<unsigned int> ids({(py::ssize_t)nq, (py::ssize_t)k});
  py::array_t<float> dists({(py::ssize_t)nq, (py::ssize_t)k});
  unsigned int* ip = ids.mutable_data();
  float* dp = dists.mutable_data();
  {
    py::gil_scoped_release release;
    f(ip, dp);
  }
  return std::make_pair(ids, dists);
}

template <class T, int METRIC>
static void add_variant(py::module_& m, const std::string& sfx) {
  struct Prefilter : wsann::PrefilterIndex { using wsann::PrefilterIndex::PrefilterIndex; };
  struct Postfilter : wsann::PostfilterVamanaIndex { using wsann::PostfilterVamanaIndex::PostfilterVamanaIndex; };
  struct Tree : wsann::VamanaRangeFilterTreeIndex { using wsann::VamanaRangeFilterTreeIndex::VamanaRangeFilterTreeIndex; };
  struct PreTree : wsann::RangeFilterTreeIndex { using wsann::RangeFilterTreeIndex::RangeFilterTreeIndex; };
  struct Super : wsann::SuperOptimizedPostfilterTree { using wsann::SuperOptimizedPostfilterTree::SuperOptimizedPostfilterTree; };

  py::class_<Prefilter>(m, ("PrefilterIndex" + sfx).c_str())
      .def(py::init([](TArray<T> points, FArray filter_values, BuildParams bp) {
             Points<T> p = check_points<T, METRIC>(points, filter_values);
             return new Prefilter(p.data, filter_values.data(), p.n, p.dim, METRIC, bp);
           }),
           "points"_a, "filter_values"_a, "build_params"_a = DEFAULT_BUILD_PARAMS)
      .def("batch_search",
           [](Prefilter& self, TArray<T> queries, FArray filters, uint64_t num_queries, QueryParams qp) {
             Batch<T> b = check_batch<T>(queries, filters, num_queries, self.dim());
             return run(b.nq, qp.k, [&](unsigned int* ids, float* dists) { self.batch_search(b.queries, b.filters, b.nq, qp, ids, dists); });
           },
           "queries"_a, "filters"_a, "num_queries"_a, "query_params"_a)
      .def("_arena_handle", [](Prefilter& self) { return (uintptr_t)self.arena().get(); })
      .def("_group_handle", [](Prefilter& self) { return (uintptr_t)self.group(); })
      // not in the reference (it persists graphs only): the finished arena in one file, and back
      .def("save_snapshot", [](Prefilter& self, const std::string& path) { self.save_snapshot(path); }, "path"_a)
      .def_static("load_snapshot", [](const std::string& path) { return new Prefilter(wsann::FromSnapshot{path}); }, "path"_a);

  py::class_<Postfilter>(m, ("PostfilterVamanaIndex" + sfx).c_str())
      .def(py::init([](TArray<T> points, FArray filters, BuildParams bp) {
             Points<T> p = check_points<T, METRIC>(points, filters);
             return new Postfilter(p.data, filters.data(), p.n, p.dim, METRIC, bp);
           }),
           "points"_a, "filters"_a, "build_params"_a = DEFAULT_BUILD_PARAMS)
      .def("batch_search",
           [](Postfilter& self, TArray<T> queries, FArray filters, uint64_t num_queries, QueryParams qp) {
             Batch<T> b = check_batch<T>(queries, filters, num_queries, self.dim());
             return run(b.nq, qp.k, [&](unsigned int* ids, float* dists) { self.batch_search(b.queries, b.filters, b.nq, qp, ids, dists); });
           },
           "queries"_a, "filters"_a, "num_queries"_a, "query_params"_a)
      .def("_arena_handle", [](Postfilter& self) { return (uintptr_t)self.arena().get(); })
      .def("_group_handle", [](Postfilter& self) { return (uintptr_t)self.group(); })
      // not in the reference (it persists graphs only): the finished arena in one file, and back
      .def("save_snapshot", [](Postfilter& self, const std::string& path) { self.save_snapshot(path); }, "path"_a)
      .def_static("load_snapshot", [](const std::string& path) { return new Postfilter(wsann::FromSnapshot{path}); }, "path"_a);

  py::class_<Tree>(m, ("VamanaRangeFilterTreeIndex" + sfx).c_str())
      .def(py::init([](TArray<T> points, FArray filter_values, int32_t cutoff, size_t split_factor, BuildParams bp) {
             Points<T> p = check_points<T, METRIC>(points, filter_values);
             return new Tree(p.data, filter_values.data(), p.n, p.dim, METRIC, cutoff, split_factor, bp);
           }),
           "points"_a, "filter_values"_a, "cutoff"_a = 1000, "split_factor"_a = 2,
           "build_params"_a = DEFAULT_BUILD_PARAMS)
      .def("batch_search",
           [](Tree& self, TArray<T> queries, FArray filters, uint64_t num_queries, const std::string& query_method,
              QueryParams qp) {
             Batch<T> b = check_batch<T>(queries, filters, num_queries, self.dim());
             return run(b.nq, qp.k, [&](unsigned int* ids, float* dists) {
               self.batch_search(b.queries, b.filters, b.nq, query_method, qp, ids, dists);
             });
           },
           "queries"_a, "filters"_a, "num_queries"_a, "query_method"_a, "query_params"_a)
      .def("_arena_handle", [](Tree& self) { return (uintptr_t)self.arena().get(); })
      .def("_group_handle", [](Tree& self) { return (uintptr_t)self.group(); })
      // not in the reference (it persists graphs only): the finished arena in one file, and back
      .def("save_snapshot", [](Tree& self, const std::string& path) { self.save_snapshot(path); }, "path"_a)
      .def_static("load_snapshot", [](const std::string& path) { return new Tree(wsann::FromSnapshot{path}); }, "path"_a)
      .def("_bucket_offsets", [](Tree& self) { return self.bucket_offsets(); });

  // python_bindings.cpp:119-127 — the tree over PrefilterIndex sub-indices
  py::class_<PreTree>(m, ("RangeFilterTreeIndex" + sfx).c_str())
      .def(py::init([](TArray<T> points, FArray filter_values, int32_t cutoff, size_t split_factor, BuildParams bp) {
             Points<T> p = check_points<T, METRIC>(points, filter_values);
             return new PreTree(p.data, filter_values.data(), p.n, p.dim, METRIC, cutoff, split_factor, bp);
           }),
           "points"_a, "filter_values"_a, "cutoff"_a = 1000, "split_factor"_a = 2,
           "build_params"_a = DEFAULT_BUILD_PARAMS)
      .def("batch_search",
           [](PreTree& self, TArray<T> queries, FArray filters, uint64_t num_queries, const std::string& query_method,
              QueryParams qp) {
             Batch<T> b = check_batch<T>(queries, filters, num_queries, self.dim());
             return run(b.nq, qp.k, [&](unsigned int* ids, float* dists) {
               self.batch_search(b.queries, b.filters, b.nq, query_method, qp, ids, dists);
             });
           },
           "queries"_a, "filters"_a, "num_queries"_a, "query_method"_a, "query_params"_a)
      .def("_arena_handle", [](PreTree& self) { return (uintptr_t)self.arena().get(); })
      .def("_group_handle", [](PreTree& self) { return (uintptr_t)self.group(); })
      // not in the reference (it persists graphs only): the finished arena in one file, and back
      .def("save_snapshot", [](PreTree& self, const std::string& path) { self.save_snapshot(path); }, "path"_a)
      .def_static("load_snapshot", [](const std::string& path) { return new PreTree(wsann::FromSnapshot{path}); }, "path"_a)
      .def("_bucket_offsets", [](PreTree& self) { return self.bucket_offsets(); });

  py::class_<Super>(m, ("SuperOptimizedPostfilterTreeIndex" + sfx).c_str())
      .def(py::init([](TArray<T> points, FArray filter_values, int32_t cutoff, float split_factor, float shift_factor,
                       BuildParams bp) {
             Points<T> p = check_points<T, METRIC>(points, filter_values);
             return new Super(p.data, filter_values.data(), p.n, p.dim, METRIC, cutoff, split_factor, shift_factor, bp);
           }),
           "points"_a, "filter_values"_a, "cutoff"_a = 1000, "split_factor"_a = 2, "shift_factor"_a = 0.5,
           "build_params"_a = DEFAULT_BUILD_PARAMS)
      .def("batch_search",
           [](Super& self, TArray<T> queries, FArray filters, uint64_t num_queries, QueryParams qp) {
             Batch<T> b = check_batch<T>(queries, filters, num_queries, self.dim());
             return run(b.nq, qp.k, [&](unsigned int* ids, float* dists) { self.batch_search(b.queries, b.filters, b.nq, qp, ids, dists); });
           },
           "queries"_a, "filters"_a, "num_queries"_a, "query_params"_a)
      .def("_arena_handle", [](Super& self) { return (uintptr_t)self.arena().get(); })
      .def("_group_handle", [](Super& self) { return (uintptr_t)self.group(); })
      // not in the reference (it persists graphs only): the finished arena in one file, and back
      .def("save_snapshot", [](Super& self, const std::string& path) { self.save_snapshot(path); }, "path"_a)
      .def_static("load_snapshot", [](const std::string& path) { return new Super(wsann::FromSnapshot{path}); }, "path"_a);

}

// The extension itself is named _window_ann_b200 so that it can live in one interpreter next to
// the reference's `window_ann` extension (CPython caches extension modules by name); the
// drop-in name is provided by rangefilteredann_b200/window_ann.py, which re-exports it.
PYBIND11_MODULE(_window_ann_b200, m) {
  m.doc() = "WindowANN Python bindings — B200-native window-search engine (drop-in for the reference module)";
  m.attr("__version__") = "b200-dev";
  m.attr("__engine__") = "wsann_cuda";

  py::module_ default_values = m.def_submodule("defaults");  // python_bindings.cpp:170-175
  default_values.attr("METRIC") = "Euclidian";
  default_values.attr("ALPHA") = 1.2;
  default_values.attr("GRAPH_DEGREE") = 64;
  default_values.attr("BEAMWIDTH") = 128;

  py::class_<QueryParams>(m, "QueryParams")
      .def(py::init<long, long, double, long, long, long, long, std::optional<float>, bool>(), "k"_a, "beam_width"_a,
           "cut"_a, "limit"_a, "degree_limit"_a, "final_beam_multiply"_a, "postfiltering_max_beam"_a,
           "min_query_to_bucket_ratio"_a, "verbose"_a);

  py::class_<BuildParams>(m, "BuildParams")
      .def(py::init<long, long, double, std::string>(), "max_degree"_a, "limit"_a, "alpha"_a, "cache_path"_a);

  add_variant<float, WS_METRIC_L2>(m, "FloatEuclidian");
  add_variant<float, WS_METRIC_MIPS>(m, "FloatMips");
  // python_bindings.cpp:74-86,233-236 (registered as "UInt8…"; wrapper.py:244-330 asks for "Uint8…" and so
  // never reaches them, SURVEY.md §A-10 — both spellings are not provided, only the registered one)
  add_variant<uint8_t, WS_METRIC_L2>(m, "UInt8Euclidian");
  add_variant<uint8_t, WS_METRIC_MIPS>(m, "UInt8Mips");
  add_variant<int8_t, WS_METRIC_L2>(m, "Int8Euclidian");
  add_variant<int8_t, WS_METRIC_MIPS>(m, "Int8Mips");

  m.def("device_count", []() {
    int c = 0;
    ws_device_count(&c);
    return c;
  });
  m.def("abi_version", []() { return ws_abi_version(); });
}
