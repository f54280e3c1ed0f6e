// optix_extend_b200.cpp -- the reference's pybind plugin class `optix.optix_mesh` (optix_extend.cpp:6-83) as a compiled
// torch extension over libdrt_b200's C ABI: same four methods, same tensor arguments, no OptiX headers.  This is the
// binding INTEGRATION.md section 4 shows; it is built in-tree by drt_b200/build.py:build_torch_plugin() and exercised by
// tests/test_gpu_plugin_ext.py.  The in-tree Python host side binds the same ABI with ctypes (drt_b200/optix.py).
#include <torch/extension.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "drt_b200.h"

class optix_mesh {
    drt_bvh* h = nullptr;
    int device = 0;

    void check(const torch::Tensor& t, const char* name, int64_t cols, c10::ScalarType dtype) const
    {
        TORCH_CHECK(t.is_cuda() && t.get_device() == device, name, " must live on cuda:", device);
        TORCH_CHECK(t.dim() == 2 && t.size(1) == cols, name, " must be [n,", cols, "]");  // assert(size(1)==..), optix_extend.cpp:17-18,25,31
        TORCH_CHECK(t.scalar_type() == dtype, name, " has the wrong dtype");
    }

   public:
    explicit optix_mesh(unsigned dev) : device((int)dev)  // optix_extend.cpp:8-12
    {
        TORCH_CHECK(drt_bvh_create((int)dev, &h) == DRT_OK, drt_last_error());
    }
    ~optix_mesh() { drt_bvh_destroy(h); }
    optix_mesh(const optix_mesh&) = delete;
    optix_mesh& operator=(const optix_mesh&) = delete;

    void update_mesh(torch::Tensor F, torch::Tensor V)  // optix_extend.cpp:14-21
    {
        check(F, "F", 3, torch::kInt32);
        check(V, "V", 3, torch::kFloat32);
        F = F.contiguous();
        V = V.contiguous();
        c10::cuda::CUDAGuard g(device);
        TORCH_CHECK(drt_bvh_build(h, F.data_ptr<int>(), (int)F.size(0), V.data_ptr<float>(), (int)V.size(0),
                                  at::cuda::getCurrentCUDAStream(device).stream()) == DRT_OK, drt_last_error());
    }

    void update_vert(torch::Tensor V)  // optix_extend.cpp:23-27 (full rebuild, like the reference)
    {
        check(V, "V", 3, torch::kFloat32);
        V = V.contiguous();
        c10::cuda::CUDAGuard g(device);
        TORCH_CHECK(drt_bvh_update_vert(h, V.data_ptr<float>(), nullptr, (int)V.size(0), /*refit=*/0,
                                        at::cuda::getCurrentCUDAStream(device).stream()) == DRT_OK, drt_last_error());
    }

    std::vector<at::Tensor> intersect(torch::Tensor Ray)  // optix_extend.cpp:29-57
    {
        check(Ray, "Ray", 6, torch::kFloat32);
        Ray = Ray.contiguous();
        c10::cuda::CUDAGuard g(device);
        auto Hit = torch::empty({Ray.size(0), 2}, Ray.options());  // {float t; int id} records, optix_extend.cpp:36-41
        auto HitI = Hit.view(torch::kInt32);
        TORCH_CHECK(drt_closest_hit(h, Ray.data_ptr<float>(), Ray.size(0), Hit.data_ptr<float>(), HitI.data_ptr<int>() + 1, 2, 2,
                                    at::cuda::getCurrentCUDAStream(device).stream()) == DRT_OK, drt_last_error());
        return {Hit.select(1, 0), HitI.select(1, 1)};  // owning views (the reference's ID is a non-owning from_blob alias, :52-54)
    }
};

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m)
{
    py::class_<optix_mesh>(m, "optix_mesh")
        .def(py::init<unsigned>())
        .def("update_mesh", &optix_mesh::update_mesh)
        .def("update_vert", &optix_mesh::update_vert)
        .def("intersect", &optix_mesh::intersect);
}
