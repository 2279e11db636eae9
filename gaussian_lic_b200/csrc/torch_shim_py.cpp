// torch_shim_py.cpp -- pybind11 face of the LibTorch shim so pytest can drive the exact C++ symbols
// a Gaussian-LIC build would link against (declarations mirror the reference headers
// rasterizer/rasterize_points.h:25-96, fused-ssim/ssim.h:7-26, simple-knn/spatial.h:14).
#include <torch/extension.h>
#include <tuple>

std::tuple<int, int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors,
                       const torch::Tensor& opacity, const torch::Tensor& scales, const torch::Tensor& rotations,
                       const float scale_modifier, const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix,
                       const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy, const int image_height,
                       const int image_width, const float limx_neg, const float limx_pos, const float limy_neg,
                       const float limy_pos, const torch::Tensor& dc, const torch::Tensor& sh, const int degree,
                       const torch::Tensor& campos, const bool prefiltered, const bool debug, const bool no_color = false);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii,
                               const torch::Tensor& colors, const torch::Tensor& scales, const torch::Tensor& rotations,
                               const float scale_modifier, const torch::Tensor& cov3D_precomp,
                               const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx,
                               const float tan_fovy, const float limx_neg, const float limx_pos, const float limy_neg,
                               const float limy_pos, const torch::Tensor& dL_dout_color, const torch::Tensor& dc,
                               const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                               const torch::Tensor& geomBuffer, const int R, const torch::Tensor& binningBuffer,
                               const torch::Tensor& imageBuffer, const int B, const torch::Tensor& sampleBuffer,
                               const float lambda_erank, const bool debug);

void adamUpdate(torch::Tensor& param, torch::Tensor& param_grad, torch::Tensor& exp_avg, torch::Tensor& exp_avg_sq,
                torch::Tensor& visible, const float lr, const float b1, const float b2, const float eps, const uint32_t N,
                const uint32_t M);

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
fusedssim(float C1, float C2, torch::Tensor& img1, torch::Tensor& img2, bool train);

torch::Tensor fusedssim_backward(float C1, float C2, torch::Tensor& img1, torch::Tensor& img2, torch::Tensor& dL_dmap,
                                 torch::Tensor& dm_dmu1, torch::Tensor& dm_dsigma1_sq, torch::Tensor& dm_dsigma12);

torch::Tensor distCUDA2(const torch::Tensor& points);

void glic_bind_host(pybind11::module_& m);   // torch_host.cpp: C++ host of the mapping-iteration body on these symbols

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    glic_bind_host(m);
    m.def("RasterizeGaussiansCUDA", &RasterizeGaussiansCUDA);
    m.def("RasterizeGaussiansBackwardCUDA", &RasterizeGaussiansBackwardCUDA);
    m.def("adamUpdate", &adamUpdate);
    m.def("fusedssim", &fusedssim);
    m.def("fusedssim_backward", &fusedssim_backward);
    m.def("distCUDA2", &distCUDA2);
}
