/* Hand-written stand-in for the header GTSAM's CMake would generate from
 * gtsam/config.h.in (reference: gtsam/config.h.in:1-95).  Test infrastructure
 * only: used by oracle/Makefile to compile the UNMODIFIED reference sources
 * where they lie under /root/reference into oracle/_ref/.  Flags match
 * Rot3=matrix, Pose3/Rot3 EXPMAP, cheirality throws,
 * fast BetweenFactor Jacobian, no TBB, STL allocator, vendored Eigen+METIS. */
#pragma once
#define GTSAM_VERSION_MAJOR 4
#define GTSAM_VERSION_MINOR 3
#define GTSAM_VERSION_PATCH 0
#define GTSAM_VERSION_NUMERIC 40300
#define GTSAM_VERSION_STRING "4.3a0"
#define GTSAM_SOURCE_TREE_DATASET_DIR "/root/reference/examples/Data"
#define GTSAM_INSTALLED_DATASET_DIR "/root/reference/examples/Data"
#define GTSAM_POSE3_EXPMAP
#define GTSAM_ROT3_EXPMAP
#define GTSAM_DT_MERGING
#define GTSAM_EIGEN_VERSION_WORLD 3
#define GTSAM_EIGEN_VERSION_MAJOR 4
#define GTSAM_EIGEN_VERSION_MINOR 0
#define GTSAM_ALLOCATOR_STL
#define GTSAM_THROW_CHEIRALITY_EXCEPTION
#define GTSAM_SUPPORT_NESTED_DISSECTION
#define GTSAM_TANGENT_PREINTEGRATION
