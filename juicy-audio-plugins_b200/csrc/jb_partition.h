// jb_partition.h -- disjoint SM partitions (CUDA green contexts) with a stream each; see jb_partition.cpp.
#pragma once

namespace jb {

struct SmPartitions {
    static constexpr int kMax = 8;
    int count = 0;          // partitions created (0: none -- the caller uses ordinary streams)
    int smsPerGroup = 0;
    void* contexts[kMax] = {};
    void* streams[kMax] = {}; // cudaStream_t, one inside each partition
    int create(int device, int groups);
    void release();
    ~SmPartitions();
    SmPartitions() = default;
    SmPartitions(const SmPartitions&) = delete;
    SmPartitions& operator=(const SmPartitions&) = delete;
};

} // namespace jb
