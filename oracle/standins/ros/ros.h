// TEST INFRASTRUCTURE ONLY -- empty stand-in for <ros/ros.h>.  src/bgkloctomap/bgkloctomap.cpp:2 includes it but uses
// nothing from it; ROS is not installed in this image.
#pragma once
