// Stand-in for nav2_costmap_2d::Costmap2D / Costmap2DROS: a raw 0..254 grid plus the robot footprint.
#pragma once
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include "geometry_msgs/msg/types.hpp"
namespace nav2_costmap_2d {
class Costmap2D {
public:
  Costmap2D(unsigned w, unsigned h, double res, double ox, double oy)
  : w_(w), h_(h), res_(res), ox_(ox), oy_(oy), cells_((size_t)w * h, 0) {}
  unsigned char * getCharMap() { return cells_.data(); }
  unsigned getSizeInCellsX() const { return w_; }
  unsigned getSizeInCellsY() const { return h_; }
  double getResolution() const { return res_; }
  double getOriginX() const { return ox_; }
  double getOriginY() const { return oy_; }
  typedef std::recursive_mutex mutex_t;                      // as in nav2_costmap_2d::Costmap2D
  mutex_t * getMutex() { return &mutex_; }
private:
  unsigned w_, h_; double res_, ox_, oy_; std::vector<unsigned char> cells_; mutex_t mutex_;
};
class Costmap2DROS {
public:
  Costmap2DROS(std::shared_ptr<Costmap2D> c, std::vector<geometry_msgs::msg::Point> fp) : c_(c), fp_(fp) {}
  Costmap2D * getCostmap() { return c_.get(); }
  std::vector<geometry_msgs::msg::Point> getRobotFootprint() const { return fp_; }
  std::string getBaseFrameID() const { return "base_link"; }
  std::string getGlobalFrameID() const { return "odom"; }
private:
  std::shared_ptr<Costmap2D> c_; std::vector<geometry_msgs::msg::Point> fp_;
};
}
