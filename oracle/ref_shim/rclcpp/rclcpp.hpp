// ORACLE shim
#pragma once
#include "rclcpp/clock.hpp"
#include "rclcpp/logging.hpp"
#include "rclcpp/node.hpp"
#include "rclcpp/time.hpp"
