// ORACLE shim: logging macros compile to nothing
#pragma once
#define RCLCPP_WARN(...) ((void)0)
#define RCLCPP_ERROR(...) ((void)0)
#define RCLCPP_INFO(...) ((void)0)
#define RCLCPP_DEBUG(...) ((void)0)
