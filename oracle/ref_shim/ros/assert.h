// ROS logging / assertion macros: assertions stay assertions, log streams compile to nothing (their arguments are not evaluated)
#pragma once
#include <cassert>
#include <cstdlib>
#define ROS_ASSERT(cond) assert(cond)
#define ROS_ASSERT_MSG(cond, ...) assert(cond)
#define ROS_BREAK() abort()
#define ROS_DEBUG_STREAM(x) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_WARN_STREAM(x) do { } while (0)
#define ROS_ERROR_STREAM(x) do { } while (0)
#define ROS_FATAL_STREAM(x) do { } while (0)
#define ROS_DEBUG(...) do { } while (0)
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
#define ROS_FATAL(...) do { } while (0)
