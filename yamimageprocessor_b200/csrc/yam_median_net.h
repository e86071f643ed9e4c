// Median selection networks (9 and 25 inputs) built from min/max exchanges.
// Host+device so tests/test_median_network.py can check them exhaustively on 0/1 inputs
// (zero-one principle) without a GPU.
#pragma once
#ifdef __CUDACC__
#define YAM_HD __host__ __device__ __forceinline__
#else
#define YAM_HD inline
#endif

template <typename V>
YAM_HD void cex(V& a, V& b) {
    V lo = a < b ? a : b;
    V hi = a < b ? b : a;
    a = lo;
    b = hi;
}

template <typename V>
YAM_HD V median9(V* p) {
    cex(p[1], p[2]); cex(p[4], p[5]); cex(p[7], p[8]); cex(p[0], p[1]); cex(p[3], p[4]); cex(p[6], p[7]);
    cex(p[1], p[2]); cex(p[4], p[5]); cex(p[7], p[8]); cex(p[0], p[3]); cex(p[5], p[8]); cex(p[4], p[7]);
    cex(p[3], p[6]); cex(p[1], p[4]); cex(p[2], p[5]); cex(p[4], p[7]); cex(p[4], p[2]); cex(p[6], p[4]);
    cex(p[4], p[2]);
    return p[4];
}

template <typename V>
YAM_HD V median25(V* p) {
    cex(p[0], p[1]); cex(p[3], p[4]); cex(p[2], p[4]); cex(p[2], p[3]); cex(p[6], p[7]); cex(p[5], p[7]);
    cex(p[5], p[6]); cex(p[9], p[10]); cex(p[8], p[10]); cex(p[8], p[9]); cex(p[12], p[13]); cex(p[11], p[13]);
    cex(p[11], p[12]); cex(p[15], p[16]); cex(p[14], p[16]); cex(p[14], p[15]); cex(p[18], p[19]); cex(p[17], p[19]);
    cex(p[17], p[18]); cex(p[21], p[22]); cex(p[20], p[22]); cex(p[20], p[21]); cex(p[23], p[24]); cex(p[2], p[5]);
    cex(p[3], p[6]); cex(p[0], p[6]); cex(p[0], p[3]); cex(p[4], p[7]); cex(p[1], p[7]); cex(p[1], p[4]);
    cex(p[11], p[14]); cex(p[8], p[14]); cex(p[8], p[11]); cex(p[12], p[15]); cex(p[9], p[15]); cex(p[9], p[12]);
    cex(p[13], p[16]); cex(p[10], p[16]); cex(p[10], p[13]); cex(p[20], p[23]); cex(p[17], p[23]); cex(p[17], p[20]);
    cex(p[21], p[24]); cex(p[18], p[24]); cex(p[18], p[21]); cex(p[19], p[22]); cex(p[8], p[17]); cex(p[9], p[18]);
    cex(p[0], p[18]); cex(p[0], p[9]); cex(p[10], p[19]); cex(p[1], p[19]); cex(p[1], p[10]); cex(p[11], p[20]);
    cex(p[2], p[20]); cex(p[2], p[11]); cex(p[12], p[21]); cex(p[3], p[21]); cex(p[3], p[12]); cex(p[13], p[22]);
    cex(p[4], p[22]); cex(p[4], p[13]); cex(p[14], p[23]); cex(p[5], p[23]); cex(p[5], p[14]); cex(p[15], p[24]);
    cex(p[6], p[24]); cex(p[6], p[15]); cex(p[7], p[16]); cex(p[7], p[19]); cex(p[13], p[21]); cex(p[15], p[23]);
    cex(p[7], p[13]); cex(p[7], p[15]); cex(p[1], p[9]); cex(p[3], p[11]); cex(p[5], p[17]); cex(p[11], p[17]);
    cex(p[9], p[17]); cex(p[4], p[10]); cex(p[6], p[12]); cex(p[7], p[14]); cex(p[4], p[6]); cex(p[4], p[7]);
    cex(p[12], p[14]); cex(p[10], p[14]); cex(p[6], p[7]); cex(p[10], p[12]); cex(p[6], p[10]); cex(p[6], p[17]);
    cex(p[12], p[17]); cex(p[7], p[17]); cex(p[7], p[10]); cex(p[12], p[18]); cex(p[7], p[12]); cex(p[10], p[18]);
    cex(p[12], p[20]); cex(p[10], p[20]); cex(p[10], p[12]);
    return p[12];
}

