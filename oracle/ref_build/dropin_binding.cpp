// Drop-in proof (test infrastructure): the reference's UNCHANGED host-side code -- the autograd Function of
// src/rasterizer/rasterizer.cpp, FusedSSIMMap / l1_loss of src/loss_utils.h, SparseGaussianAdam of
// src/optim_utils.h -- compiled from /root/reference and linked against OUR libraries
// (glic_b200_torch.so + libglic_b200.so) instead of the reference's rasterize_points.cu / ssim.cu / adam.cu.
// If any of the six boundary symbols were missing or had a different signature, this would not link.
#include <torch/extension.h>
#include <memory>
#include <vector>

#include "rasterizer/rasterizer.h"   // reference
#include "loss_utils.h"              // reference (header only)
#include "optim_utils.h"             // reference (header only)
#include "simple-knn/spatial.h"      // reference declaration of distCUDA2

namespace {
using torch::Tensor;

std::vector<Tensor> autograd_rasterize(Tensor means3D, Tensor means2D, Tensor opacities, Tensor dc, Tensor shs, Tensor scales,
                                       Tensor rotations, Tensor bg, Tensor viewmatrix, Tensor projmatrix, Tensor campos,
                                       int64_t H, int64_t W, double tanfovx, double tanfovy, double limx_neg, double limx_pos,
                                       double limy_neg, double limy_pos, int64_t sh_degree, bool no_color, double lambda_erank) {
    GaussianRasterizationSettings settings((int)H, (int)W, (float)tanfovx, (float)tanfovy, (float)limx_neg, (float)limx_pos,
                                           (float)limy_neg, (float)limy_pos, bg, 1.0f, viewmatrix, projmatrix, (int)sh_degree,
                                           campos, false, false, no_color, (float)lambda_erank);
    GaussianRasterizer rasterizer(settings);
    Tensor colors_precomp, cov3D_precomp;
    auto r = rasterizer.forward(means3D, means2D, opacities, dc, shs, colors_precomp, scales, rotations, cov3D_precomp);
    return {std::get<0>(r), std::get<1>(r), std::get<2>(r)};
}

Tensor fused_ssim_autograd(Tensor img1, Tensor img2) { return loss_utils::fused_ssim(img1, img2); }
Tensor l1_autograd(Tensor a, Tensor b) { return loss_utils::l1_loss(a, b); }

// One SparseGaussianAdam step over `params` (each with .grad set), like gaussian.cpp:703-707.
void sparse_adam_step(std::vector<Tensor> params, std::vector<double> lrs, Tensor visible, int64_t N) {
    std::unique_ptr<SparseGaussianAdam> opt;
    for (size_t i = 0; i < params.size(); ++i) {
        std::vector<Tensor> one{params[i]};
        if (i == 0) opt.reset(new SparseGaussianAdam(one, 0.0, 1e-15));
        else opt->add_param_group(one);
        opt->param_groups()[i].options().set_lr(lrs[i]);
    }
    opt->set_visibility_and_N(visible, N);
    opt->step();
}
}  // namespace

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("autograd_rasterize", &autograd_rasterize);
    m.def("fused_ssim_autograd", &fused_ssim_autograd);
    m.def("l1_autograd", &l1_autograd);
    m.def("sparse_adam_step", &sparse_adam_step);
    m.def("distCUDA2", &distCUDA2);
}
