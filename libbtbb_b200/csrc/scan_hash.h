/* scan_hash.h -- hash functions shared by the table builder (host) and the scan kernels. */
#ifndef BTBB_B200_SCAN_HASH_H
#define BTBB_B200_SCAN_HASH_H
#include "bt_math.h"

BT_HD uint32_t bt_bloom_h1(uint32_t s32, int log2bits) { return (s32 * 0x9E3779B1u) >> (32 - log2bits); }
BT_HD uint32_t bt_bloom_h2(uint32_t s32, int log2bits) { return (s32 * 0x85EBCA6Bu + 0x27D4EB2Fu) >> (32 - log2bits); }
BT_HD uint64_t bt_err_hash(uint64_t syn, int log2cap) { return (syn * 0x9E3779B97F4A7C15ULL) >> (64 - log2cap); }

#endif
