// Python binding for the UNMODIFIED reference hot path ("Oracle A").
//
// TEST / BENCH INFRASTRUCTURE ONLY.  This translation unit is compiled together
// with the reference's own sources where they lie under /root/reference
// (src/rasterizer/**, src/fused-ssim/ssim.cu, src/simple-knn/*.cu) by
// oracle/ref_build/build_ref.py into oracle/_ref/glic_ref_ext.so.  No reference
// source is copied into this repository.  The product (gaussian_lic_b200/) never
// imports it.
//
// Besides the six boundary functions it exposes the reference's
// own autograd op (src/rasterizer/rasterizer.cpp) and helpers that slice the
// reference's opaque byte arenas with the reference's own `fromChunk` carvers so
// tests can compare tile lists / ranges bit-for-bit.
#include <torch/extension.h>
#include <memory>
#include <tuple>
#include <vector>

#include "rasterizer/rasterizer.h"        // GaussianRasterizerFunction, settings (reference)
#include "rasterizer/rasterize_points.h"  // boundary functions (reference)
#include "fused-ssim/ssim.h"              // fusedssim / fusedssim_backward (reference)
#include "simple-knn/spatial.h"           // distCUDA2 (reference)
#include "loss_utils.h"                   // FusedSSIMMap, l1_loss (reference, header only)
#include "optim_utils.h"                  // SparseGaussianAdam (reference, header only)
#include "rasterizer/cuda_rasterizer/config.h"
#include "rasterizer/cuda_rasterizer/rasterizer_impl.h"

namespace {

using torch::Tensor;

Tensor from_dev(const void* p, std::vector<int64_t> shape, torch::ScalarType dt, const Tensor& like) {
    auto opts = torch::TensorOptions().dtype(dt).device(like.device());
    return torch::from_blob(const_cast<void*>(p), shape, opts).clone();
}

// geomBuffer -> (depths[P], means2D[P,2], conic_opacity[P,4], rgb[P,3], tiles_touched[P], point_offsets[P], clamped[P,3])
std::vector<Tensor> slice_geom(const Tensor& geom, int64_t P) {
    char* chunk = reinterpret_cast<char*>(geom.data_ptr());
    auto g = CudaRasterizer::GeometryState::fromChunk(chunk, (size_t)P);
    return {
        from_dev(g.depths, {P}, torch::kFloat32, geom),
        from_dev(g.means2D, {P, 2}, torch::kFloat32, geom),
        from_dev(g.conic_opacity, {P, 4}, torch::kFloat32, geom),
        from_dev(g.rgb, {P, 3}, torch::kFloat32, geom),
        from_dev(g.tiles_touched, {P}, torch::kInt32, geom),
        from_dev(g.point_offsets, {P}, torch::kInt32, geom),
        from_dev(g.clamped, {P, 3}, torch::kBool, geom),
        from_dev(g.cov3D, {P, 6}, torch::kFloat32, geom),
    };
}

// binningBuffer -> (point_list[R] i32, keys_sorted[R] i64)
std::vector<Tensor> slice_binning(const Tensor& binning, int64_t R) {
    if (R == 0) {
        return {torch::empty({0}, torch::TensorOptions().dtype(torch::kInt32).device(binning.device())),
                torch::empty({0}, torch::TensorOptions().dtype(torch::kInt64).device(binning.device()))};
    }
    char* chunk = reinterpret_cast<char*>(binning.data_ptr());
    auto b = CudaRasterizer::BinningState::fromChunk(chunk, (size_t)R);
    return {
        from_dev(b.point_list, {R}, torch::kInt32, binning),
        from_dev(b.point_list_keys, {R}, torch::kInt64, binning),
    };
}

// imgBuffer -> (ranges[T,2] i32, n_contrib[H*W] i32, max_contrib[T] i32, bucket_offsets[T] i32)
std::vector<Tensor> slice_image(const Tensor& img, int64_t H, int64_t W) {
    int64_t tx = (W + BLOCK_X - 1) / BLOCK_X, ty = (H + BLOCK_Y - 1) / BLOCK_Y;
    int64_t T = tx * ty;
    char* chunk = reinterpret_cast<char*>(img.data_ptr());
    auto s = CudaRasterizer::ImageState::fromChunk(chunk, (size_t)(H * W), (size_t)T);
    return {
        from_dev(s.ranges, {T, 2}, torch::kInt32, img),
        from_dev(s.n_contrib, {H * W}, torch::kInt32, img),
        from_dev(s.max_contrib, {T}, torch::kInt32, img),
        from_dev(s.bucket_offsets, {T}, torch::kInt32, img),
    };
}

// The reference's autograd op driven exactly like src/rasterizer/renderer.cpp:21-88
// minus the Camera/GaussianModel classes (they pull in ROS/Eigen/OpenCV): the caller
// passes already-activated opacity/scales/rotations.
std::vector<Tensor> autograd_rasterize(
    Tensor means3D, Tensor means2D, Tensor opacities, Tensor dc, Tensor shs, Tensor scales, Tensor rotations,
    Tensor bg, Tensor viewmatrix, Tensor projmatrix, Tensor campos,
    int64_t H, int64_t W, double tanfovx, double tanfovy,
    double limx_neg, double limx_pos, double limy_neg, double limy_pos,
    int64_t sh_degree, bool no_color, double lambda_erank) {
    GaussianRasterizationSettings settings(
        (int)H, (int)W, (float)tanfovx, (float)tanfovy,
        (float)limx_neg, (float)limx_pos, (float)limy_neg, (float)limy_pos,
        bg, 1.0f, viewmatrix, projmatrix, (int)sh_degree, campos,
        /*prefiltered=*/false, /*debug=*/false, no_color, (float)lambda_erank);
    GaussianRasterizer rasterizer(settings);
    Tensor colors_precomp, cov3D_precomp;
    auto r = rasterizer.forward(means3D, means2D, opacities, dc, shs, colors_precomp, scales, rotations, cov3D_precomp);
    return {std::get<0>(r), std::get<1>(r), std::get<2>(r)};
}

// One step of the reference's SparseGaussianAdam over `params` (each with .grad set), like gaussian.cpp:399-424,703-707.
// The optimiser object is kept alive across calls (Adam state is keyed by TensorImpl*).
void sparse_adam_step(std::vector<Tensor> params, std::vector<double> lrs, Tensor visible, int64_t N) {
    static std::unique_ptr<SparseGaussianAdam> opt;
    static std::vector<void*> owners;
    std::vector<void*> now;
    for (auto& p : params) now.push_back(p.unsafeGetTensorImpl());
    if (!opt || now != owners) {
        for (size_t i = 0; i < params.size(); ++i) {
            std::vector<Tensor> one{params[i]};
            if (i == 0) opt.reset(new SparseGaussianAdam(one, 0.0, 1e-15));
            else opt->add_param_group(one);
            opt->param_groups()[i].options().set_lr(lrs[i]);
        }
        owners = now;
    }
    opt->set_visibility_and_N(visible, N);
    opt->step();
}

extern "C" int ref_cub_sort_pairs(void* temp, size_t* temp_bytes, const uint64_t* keys_in, uint64_t* keys_out,
                                  const uint32_t* vals_in, uint32_t* vals_out, int64_t n, int end_bit, void* stream);

// cub::DeviceRadixSort::SortPairs exactly as rasterizer_impl.cu:419-424 calls it; keys int64-viewed u64, values int32-viewed u32.
// `temp` may be an empty tensor: then only the required size is returned.
int64_t cub_sort_pairs(Tensor temp, Tensor keys_in, Tensor keys_out, Tensor vals_in, Tensor vals_out, int64_t end_bit) {
    size_t bytes = (size_t)temp.numel();
    const int err = ref_cub_sort_pairs(temp.numel() ? temp.data_ptr() : nullptr, &bytes,
                                       reinterpret_cast<const uint64_t*>(keys_in.data_ptr()), reinterpret_cast<uint64_t*>(keys_out.data_ptr()),
                                       reinterpret_cast<const uint32_t*>(vals_in.data_ptr()), reinterpret_cast<uint32_t*>(vals_out.data_ptr()),
                                       keys_in.numel(), (int)end_bit, nullptr);
    TORCH_CHECK(err == 0, "cub::DeviceRadixSort::SortPairs failed: ", err);
    return (int64_t)bytes;
}

Tensor fused_ssim_autograd(Tensor img1, Tensor img2) { return loss_utils::fused_ssim(img1, img2); }
Tensor l1_autograd(Tensor a, Tensor b) { return loss_utils::l1_loss(a, b); }

}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
    m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
    m.def("adamUpdate", &adamUpdate);
    m.def("fusedssim", &fusedssim);
    m.def("fusedssim_backward", &fusedssim_backward);
    m.def("distCUDA2", &distCUDA2);
    m.def("autograd_rasterize", &autograd_rasterize);
    m.def("fused_ssim_autograd", &fused_ssim_autograd);
    m.def("l1_autograd", &l1_autograd);
    m.def("sparse_adam_step", &sparse_adam_step);
    m.def("cub_sort_pairs", &cub_sort_pairs);
    m.def("slice_geom", &slice_geom);
    m.def("slice_binning", &slice_binning);
    m.def("slice_image", &slice_image);
}
