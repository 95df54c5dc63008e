// costs.cuh -- MPPICosts: host side of the MPPI running cost (API of PI/costs.cuh:59-294).
// Holds CostParams, the costmap and the world->texture transform; the device evaluation
// (computeCost and its parts, PI/costs.cu:301-414) is fused into the rollout kernel of
// libmppi_b200.so.  Controllers bound to this object pick changes up through version counters, so
// paramsToDevice() costs nothing until the next computeControl, which uploads asynchronously.
#ifndef MPPI_COSTS_CUH_
#define MPPI_COSTS_CUH_
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include <Eigen/Dense>
#include <cuda_runtime.h>

#if __has_include(<autorally_control/PathIntegralParamsConfig.h>)
#include <autorally_control/PathIntegralParamsConfig.h>
#endif

#include "../../mppi_b200.h"
#include "managed.cuh"
#include "npz_io.h"
#include "param_getter.h"

namespace autorally_control {

class MPPICosts : public Managed {
 public:
  /// Same fields, same order as the reference (PI/costs.cuh:67-85).
  typedef struct {
    float desired_speed;
    float speed_coeff;
    float track_coeff;
    float max_slip_ang;
    float slip_penalty;
    float track_slop;
    float crash_coeff;
    float steering_coeff;
    float throttle_coeff;
    float boundary_threshold;
    float discount;
    int num_timesteps;
    int grid_res;
    float3 r_c1;
    float3 r_c2;
    float3 trs;
  } CostParams;

  CostParams params_;

  MPPICosts(int width, int height) : l1_cost_(false), width_(width), height_(height) {
    zero_params();
    initCostmap();
  }

  MPPICosts(std::map<std::string, XmlRpc::XmlRpcValue> *params) : l1_cost_(false), width_(0), height_(0) {
    zero_params();
    Eigen::Matrix3f R;
    Eigen::Array3f trs;
    track_costs_ = loadTrackData((std::string)(*params)["map_path"], R, trs);
    updateTransform(R, trs);
    updateParams(params);
    costmapToTexture();
  }

  void allocateTexMem() {}  // the texture lives in each controller's context

  void updateParams_dcfg(autorally_control::PathIntegralParamsConfig config) {
    params_.desired_speed = (float)config.desired_speed;
    params_.speed_coeff = (float)config.speed_coefficient;
    params_.track_coeff = (float)config.track_coefficient;
    params_.max_slip_ang = (float)config.max_slip_angle;
    params_.slip_penalty = (float)config.slip_penalty;
    params_.crash_coeff = (float)config.crash_coefficient;
    params_.track_slop = (float)config.track_slop;
    params_.steering_coeff = (float)config.steering_coeff;
    params_.throttle_coeff = (float)config.throttle_coeff;
    paramsToDevice();
  }

  void initCostmap() {
    float4 z; z.x = z.y = z.z = z.w = 0;
    track_costs_.assign((size_t)width_ * height_, z);
  }

  /// Replace one channel of the costmap (row-major H x W) and re-publish it.
  void costmapToTexture(float *costmap, int channel = 0) {
    for (size_t i = 0; i < (size_t)width_ * height_; i++) {
      float4 &t = track_costs_[i];
      (channel == 0 ? t.x : channel == 1 ? t.y : channel == 2 ? t.z : t.w) = costmap[i];
    }
    costmapToTexture();
  }
  void costmapToTexture() { map_version_++; }

  void updateParams(std::map<std::string, XmlRpc::XmlRpcValue> *params) {
    l1_cost_ = (bool)(*params)["l1_cost"];
    params_.desired_speed = (float)(double)(*params)["desired_speed"];
    params_.speed_coeff = (float)(double)(*params)["speed_coefficient"];
    params_.track_coeff = (float)(double)(*params)["track_coefficient"];
    params_.max_slip_ang = (float)(double)(*params)["max_slip_angle"];
    params_.slip_penalty = (float)(double)(*params)["slip_penalty"];
    params_.track_slop = (float)(double)(*params)["track_slop"];
    params_.crash_coeff = (float)(double)(*params)["crash_coeff"];
    params_.steering_coeff = (float)(double)(*params)["steering_coeff"];
    params_.throttle_coeff = (float)(double)(*params)["throttle_coeff"];
    params_.boundary_threshold = (float)(double)(*params)["boundary_threshold"];
    params_.discount = (float)(double)(*params)["discount"];
    params_.num_timesteps = (int)(*params)["num_timesteps"];
    paramsToDevice();
  }

  void updateTransform(Eigen::MatrixXf m, Eigen::ArrayXf trs) {
    params_.r_c1.x = m(0, 0); params_.r_c1.y = m(1, 0); params_.r_c1.z = m(2, 0);
    params_.r_c2.x = m(0, 1); params_.r_c2.y = m(1, 1); params_.r_c2.z = m(2, 1);
    params_.trs.x = trs(0); params_.trs.y = trs(1); params_.trs.z = trs(2);
    paramsToDevice();
  }

  /// Costmap npz (keys xBounds, yBounds, pixelsPerMeter, channel0..3; PI/costs.cu:190-232).
  std::vector<float4> loadTrackData(std::string map_path, Eigen::Matrix3f &R, Eigen::Array3f &trs) {
    std::vector<float4> out;
    if (!fileExists(map_path)) {
      fprintf(stderr, "Could not load costmap at path: %s\n", map_path.c_str());
      return out;
    }
    npz::Archive d = npz::load(map_path);
    const float x_min = (float)d.at("xBounds").at(0), x_max = (float)d.at("xBounds").at(1);
    const float y_min = (float)d.at("yBounds").at(0), y_max = (float)d.at("yBounds").at(1);
    const float ppm = (float)d.at("pixelsPerMeter").at(0);
    width_ = int((x_max - x_min) * ppm);
    height_ = int((y_max - y_min) * ppm);
    initCostmap();
    out.resize((size_t)width_ * height_);
    const char *names[4] = {"channel0", "channel1", "channel2", "channel3"};
    for (int c = 0; c < 4; c++) {
      const bool have = d.count(names[c]) && d.at(names[c]).num_vals() >= out.size();
      const npz::Array *a = have ? &d.at(names[c]) : NULL;
      for (size_t i = 0; i < out.size(); i++) {
        const float v = a ? (float)a->at(i) : 0.0f;
        (c == 0 ? out[i].x : c == 1 ? out[i].y : c == 2 ? out[i].z : out[i].w) = v;
      }
    }
    R.setZero();
    R(0, 0) = 1. / (x_max - x_min); R(1, 1) = 1. / (y_max - y_min); R(2, 2) = 1;
    trs(0) = -x_min / (x_max - x_min); trs(1) = -y_min / (y_max - y_min); trs(2) = 1;
    return out;
  }

  void paramsToDevice() { params_version_++; }
  void getCostInfo() {}
  float getDesiredSpeed() { return params_.desired_speed; }
  void setDesiredSpeed(float desired_speed) { params_.desired_speed = desired_speed; paramsToDevice(); }
  void updateCostmap(std::vector<int>, std::vector<float>) {}    // empty in the reference too
  void updateObstacles(std::vector<int>, std::vector<float>) {}  // (PI/costs.cu:297-299)
  void freeCudaMem() {}

  // ---- bridge to the C ABI ----
  void setL1Cost(bool l1) { l1_cost_ = l1; paramsToDevice(); }
  int width() const { return width_; }
  int height() const { return height_; }
  unsigned long paramsVersion() const { return params_version_; }
  unsigned long mapVersion() const { return map_version_; }
  int uploadParamsTo(mppi_ctx *ctx) const {
    mppi_cost_params p;
    p.desired_speed = params_.desired_speed; p.speed_coeff = params_.speed_coeff; p.track_coeff = params_.track_coeff;
    p.max_slip_ang = params_.max_slip_ang; p.slip_penalty = params_.slip_penalty; p.track_slop = params_.track_slop;
    p.crash_coeff = params_.crash_coeff; p.steering_coeff = params_.steering_coeff; p.throttle_coeff = params_.throttle_coeff;
    p.boundary_threshold = params_.boundary_threshold; p.discount = params_.discount;
    p.num_timesteps = params_.num_timesteps; p.grid_res = params_.grid_res;
    p.r_c1[0] = params_.r_c1.x; p.r_c1[1] = params_.r_c1.y; p.r_c1[2] = params_.r_c1.z;
    p.r_c2[0] = params_.r_c2.x; p.r_c2[1] = params_.r_c2.y; p.r_c2[2] = params_.r_c2.z;
    p.trs[0] = params_.trs.x; p.trs[1] = params_.trs.y; p.trs[2] = params_.trs.z;
    p.l1_cost = l1_cost_ ? 1 : 0;
    return mppi_set_cost_params(ctx, &p);
  }
  int uploadMapTo(mppi_ctx *ctx) const {
    if (track_costs_.empty()) return MPPI_ERR_NOT_READY;
    return mppi_set_costmap(ctx, reinterpret_cast<const float *>(track_costs_.data()), width_, height_, 4);
  }

 protected:
  void zero_params() {
    params_ = CostParams();
    params_.r_c1.x = 1; params_.r_c2.y = 1; params_.trs.z = 1;
  }
  const float FRONT_D = 0.5;
  const float BACK_D = -0.5;
  bool l1_cost_;
  int width_, height_;
  std::vector<float4> track_costs_;
  unsigned long params_version_ = 0, map_version_ = 0;
};

}  // namespace autorally_control
#endif
