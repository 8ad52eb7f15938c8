// cavlc.cpp -- see cavlc.hpp.  Tables: ITU-T H.264 (03/2005) Tables 9-5, 9-7..9-10.
#include "cavlc.hpp"
#include <cstring>
#include <mutex>

namespace b200 {
namespace {

// ---- Table 9-5 coeff_token, (length, code) indexed [vlcTable][trailingOnes][totalCoeff] ----
const uint8_t kTokLen[3][4][17] = {
    {{1, 6, 8, 9, 10, 11, 13, 13, 13, 14, 14, 15, 15, 16, 16, 16, 16},
     {0, 2, 6, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 15, 16, 16, 16},
     {0, 0, 3, 7, 8, 9, 10, 11, 13, 13, 14, 14, 15, 15, 16, 16, 16},
     {0, 0, 0, 5, 6, 7, 8, 9, 10, 11, 13, 14, 14, 15, 15, 16, 16}},
    {{2, 6, 6, 7, 8, 8, 9, 11, 11, 12, 12, 12, 13, 13, 13, 14, 14},
     {0, 2, 5, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 14, 14, 14},
     {0, 0, 3, 6, 6, 7, 8, 9, 11, 11, 12, 12, 13, 13, 13, 14, 14},
     {0, 0, 0, 4, 4, 5, 6, 6, 7, 9, 11, 11, 12, 13, 13, 13, 14}},
    {{4, 6, 6, 6, 7, 7, 7, 7, 8, 8, 9, 9, 9, 10, 10, 10, 10},
     {0, 4, 5, 5, 5, 5, 6, 6, 7, 8, 8, 9, 9, 9, 10, 10, 10},
     {0, 0, 4, 5, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 10},
     {0, 0, 0, 4, 4, 4, 4, 4, 5, 6, 7, 8, 8, 9, 10, 10, 10}}};
const uint8_t kTokCode[3][4][17] = {
    {{1, 5, 7, 7, 7, 7, 15, 11, 8, 15, 11, 15, 11, 15, 11, 7, 4},
     {0, 1, 4, 6, 6, 6, 6, 14, 10, 14, 10, 14, 10, 1, 14, 10, 6},
     {0, 0, 1, 5, 5, 5, 5, 5, 13, 9, 13, 9, 13, 9, 13, 9, 5},
     {0, 0, 0, 3, 3, 4, 4, 4, 4, 4, 12, 12, 8, 12, 8, 12, 8}},
    {{3, 11, 7, 7, 7, 4, 7, 15, 11, 15, 11, 8, 15, 11, 7, 9, 7},
     {0, 2, 7, 10, 6, 6, 6, 6, 14, 10, 14, 10, 14, 10, 11, 8, 6},
     {0, 0, 3, 9, 5, 5, 5, 5, 13, 9, 13, 9, 13, 9, 6, 10, 5},
     {0, 0, 0, 5, 4, 6, 8, 4, 4, 4, 12, 8, 12, 12, 8, 1, 4}},
    {{15, 15, 11, 8, 15, 11, 9, 8, 15, 11, 15, 11, 8, 13, 9, 5, 1},
     {0, 14, 15, 12, 10, 8, 14, 10, 14, 14, 10, 14, 10, 7, 12, 8, 4},
     {0, 0, 13, 14, 11, 9, 13, 9, 13, 10, 13, 9, 13, 9, 11, 7, 3},
     {0, 0, 0, 12, 11, 10, 9, 8, 13, 12, 12, 12, 8, 12, 10, 6, 2}}};
// chroma DC (nC == -1), [trailingOnes][totalCoeff]
const uint8_t kTokDcLen[4][5] = {{2, 6, 6, 6, 6}, {0, 1, 6, 7, 8}, {0, 0, 3, 7, 8}, {0, 0, 0, 6, 7}};
const uint8_t kTokDcCode[4][5] = {{1, 7, 4, 3, 2}, {0, 1, 6, 3, 3}, {0, 0, 1, 2, 2}, {0, 0, 0, 5, 0}};

// ---- Tables 9-7/9-8 total_zeros for 4x4 blocks, [totalCoeff-1][total_zeros] ----
const uint8_t kTzLen[15][16] = {
    {1, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 9},
    {3, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 6, 6, 6, 6},
    {4, 3, 3, 3, 4, 4, 3, 3, 4, 5, 5, 6, 5, 6},
    {5, 3, 4, 4, 3, 3, 3, 4, 3, 4, 5, 5, 5},
    {4, 4, 4, 3, 3, 3, 3, 3, 4, 5, 4, 5},
    {6, 5, 3, 3, 3, 3, 3, 3, 4, 3, 6},
    {6, 5, 3, 3, 3, 2, 3, 4, 3, 6},
    {6, 4, 5, 3, 2, 2, 3, 3, 6},
    {6, 6, 4, 2, 2, 3, 2, 5},
    {5, 5, 3, 2, 2, 2, 4},
    {4, 4, 3, 3, 1, 3},
    {4, 4, 2, 1, 3},
    {3, 3, 1, 2},
    {2, 2, 1},
    {1, 1}};
const uint8_t kTzCode[15][16] = {
    {1, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 3, 2, 1},
    {7, 6, 5, 4, 3, 5, 4, 3, 2, 3, 2, 3, 2, 1, 0},
    {5, 7, 6, 5, 4, 3, 4, 3, 2, 3, 2, 1, 1, 0},
    {3, 7, 5, 4, 6, 5, 4, 3, 3, 2, 2, 1, 0},
    {5, 4, 3, 7, 6, 5, 4, 3, 2, 1, 1, 0},
    {1, 1, 7, 6, 5, 4, 3, 2, 1, 1, 0},
    {1, 1, 5, 4, 3, 3, 2, 1, 1, 0},
    {1, 1, 1, 3, 3, 2, 2, 1, 0},
    {1, 0, 1, 3, 2, 1, 1, 1},
    {1, 0, 1, 3, 2, 1, 1},
    {0, 1, 1, 2, 1, 3},
    {0, 1, 1, 1, 1},
    {0, 1, 1, 1},
    {0, 1, 1},
    {0, 1}};
// Table 9-9 total_zeros for chroma DC 2x2, [totalCoeff-1][total_zeros]
const uint8_t kTzDcLen[3][4] = {{1, 2, 3, 3}, {1, 2, 2, 0}, {1, 1, 0, 0}};
const uint8_t kTzDcCode[3][4] = {{1, 1, 1, 0}, {1, 1, 0, 0}, {1, 0, 0, 0}};
// Table 9-10 run_before, [min(zerosLeft,7)-1][run_before]
const uint8_t kRunLen[7][15] = {{1, 1},
                                {1, 2, 2},
                                {2, 2, 2, 2},
                                {2, 2, 2, 3, 3},
                                {2, 2, 3, 3, 3, 3},
                                {2, 3, 3, 3, 3, 3, 3},
                                {3, 3, 3, 3, 3, 3, 3, 4, 5, 6, 7, 8, 9, 10, 11}};
const uint8_t kRunCode[7][15] = {{1, 0},
                                 {1, 1, 0},
                                 {3, 2, 1, 0},
                                 {3, 2, 1, 1, 0},
                                 {3, 2, 3, 2, 1, 0},
                                 {3, 0, 1, 3, 2, 5, 4},
                                 {7, 6, 5, 4, 3, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1}};

// ---- expanded prefix LUTs.  entry: bits 0-4 length (0 = invalid), 5-6 trailingOnes, 7-11 totalCoeff
uint16_t gTok[3][1 << 16];
uint16_t gTokDc[1 << 8];
// entry: low nibble = length (0 invalid), high nibble = value
uint8_t gTz[15][1 << 9];
uint8_t gTzDc[3][1 << 3];
uint8_t gRun[7][1 << 11];  // value up to 14 -> (value<<4)|len  len up to 11 fits in 4 bits
std::once_flag gOnce;

void buildTables() {
    for (int t = 0; t < 3; t++) {
        std::memset(gTok[t], 0, sizeof gTok[t]);
        for (int t1 = 0; t1 < 4; t1++)
            for (int tc = 0; tc < 17; tc++) {
                int len = kTokLen[t][t1][tc];
                if (!len) continue;
                uint32_t first = (uint32_t)kTokCode[t][t1][tc] << (16 - len);
                uint16_t e = (uint16_t)(len | (t1 << 5) | (tc << 7));
                for (uint32_t k = 0; k < (1u << (16 - len)); k++) gTok[t][first + k] = e;
            }
    }
    std::memset(gTokDc, 0, sizeof gTokDc);
    for (int t1 = 0; t1 < 4; t1++)
        for (int tc = 0; tc < 5; tc++) {
            int len = kTokDcLen[t1][tc];
            if (!len) continue;
            uint32_t first = (uint32_t)kTokDcCode[t1][tc] << (8 - len);
            uint16_t e = (uint16_t)(len | (t1 << 5) | (tc << 7));
            for (uint32_t k = 0; k < (1u << (8 - len)); k++) gTokDc[first + k] = e;
        }
    std::memset(gTz, 0, sizeof gTz);
    for (int tc = 1; tc <= 15; tc++)
        for (int tz = 0; tz <= 16 - tc; tz++) {
            int len = kTzLen[tc - 1][tz];
            if (!len) continue;
            uint32_t first = (uint32_t)kTzCode[tc - 1][tz] << (9 - len);
            for (uint32_t k = 0; k < (1u << (9 - len)); k++) gTz[tc - 1][first + k] = (uint8_t)((tz << 4) | len);
        }
    std::memset(gTzDc, 0, sizeof gTzDc);
    for (int tc = 1; tc <= 3; tc++)
        for (int tz = 0; tz <= 4 - tc; tz++) {
            int len = kTzDcLen[tc - 1][tz];
            if (!len) continue;
            uint32_t first = (uint32_t)kTzDcCode[tc - 1][tz] << (3 - len);
            for (uint32_t k = 0; k < (1u << (3 - len)); k++) gTzDc[tc - 1][first + k] = (uint8_t)((tz << 4) | len);
        }
    std::memset(gRun, 0, sizeof gRun);
    for (int zl = 1; zl <= 7; zl++) {
        int maxRun = zl < 7 ? zl : 14;
        for (int r = 0; r <= maxRun; r++) {
            int len = kRunLen[zl - 1][r];
            if (!len) continue;
            uint32_t first = (uint32_t)kRunCode[zl - 1][r] << (11 - len);
            for (uint32_t k = 0; k < (1u << (11 - len)); k++) gRun[zl - 1][first + k] = (uint8_t)((r << 4) | len);
        }
    }
}

}  // namespace

void cavlcInit() { std::call_once(gOnce, buildTables); }

// test hook: decode one coeff_token for a 16-bit left-aligned prefix.  returns len | t1<<5 | tc<<7, 0 = invalid
extern "C" uint32_t b200_cavlc_probe(int kind, int index, uint32_t prefix16) {
    cavlcInit();
    switch (kind) {
        case 0:  // coeff_token, index = nC
            if (index < 0) return gTokDc[prefix16 >> 8];
            if (index >= 8) {
                uint32_t c = prefix16 >> 10;  // 6-bit FLC
                if (c == 3) return 6;         // TotalCoeff 0
                uint32_t t1 = c & 3, tc = (c >> 2) + 1;
                if (t1 > tc) return 0;
                return 6 | (t1 << 5) | (tc << 7);
            }
            return gTok[index < 2 ? 0 : index < 4 ? 1 : 2][prefix16];
        case 1:  // total_zeros 4x4, index = totalCoeff
            return gTz[index - 1][prefix16 >> 7];
        case 2:  // total_zeros chroma DC
            return gTzDc[index - 1][prefix16 >> 13];
        case 3:  // run_before, index = zerosLeft
            return gRun[(index > 7 ? 7 : index) - 1][prefix16 >> 5];
    }
    return 0;
}

CavlcResult cavlcResidualBlock(BitReader &br, int16_t *out, int nC, int maxNumCoeff) {
    CavlcResult bad{-1, 0};
    uint32_t bits = br.show32();
    uint32_t e;
    if (nC < 0) {
        e = gTokDc[bits >> 24];
    } else if (nC < 8) {
        e = gTok[nC < 2 ? 0 : nC < 4 ? 1 : 2][bits >> 16];
    } else {
        uint32_t c = bits >> 26;
        if (c == 3) {
            e = 6;
        } else {
            uint32_t t1 = c & 3, tc = (c >> 2) + 1;
            e = (t1 > tc) ? 0 : (6 | (t1 << 5) | (tc << 7));
        }
    }
    unsigned len = e & 31;
    if (!len) return bad;
    int totalCoeff = (int)(e >> 7);
    int trailingOnes = (int)((e >> 5) & 3);
    if (!br.skip(len)) return bad;
    if (totalCoeff > maxNumCoeff) return bad;
    if (totalCoeff == 0) return CavlcResult{0, 0};

    int level[16];
    int i = 0;
    if (trailingOnes) {
        uint32_t signs;
        if (!br.get((unsigned)trailingOnes, signs)) return bad;
        for (int k = trailingOnes - 1; k >= 0; k--) level[i++] = (signs >> k) & 1 ? -1 : 1;
    }
    int suffixLength = (totalCoeff > 10 && trailingOnes < 3) ? 1 : 0;
    for (; i < totalCoeff; i++) {
        uint32_t w = br.show32();
        if ((w >> 16) == 0) return bad;  // level_prefix > 15 does not exist in Baseline
        int prefix = __builtin_clz(w);
        if (!br.skip((unsigned)prefix + 1)) return bad;
        int levelCode = (prefix < 15 ? prefix : 15) << suffixLength;
        int suffixSize = suffixLength;
        if (prefix == 14 && suffixLength == 0) suffixSize = 4;
        if (prefix == 15) suffixSize = 12;
        if (suffixSize) {
            uint32_t s;
            if (!br.get((unsigned)suffixSize, s)) return bad;
            levelCode += (int)s;
        }
        if (prefix == 15 && suffixLength == 0) levelCode += 15;
        if (i == trailingOnes && trailingOnes < 3) levelCode += 2;
        int mag = (levelCode + 2) >> 1;
        if (suffixLength == 0) suffixLength = 1;
        if (mag > (3 << (suffixLength - 1)) && suffixLength < 6) suffixLength++;
        level[i] = (levelCode & 1) ? -mag : mag;
    }

    int zerosLeft = 0;
    if (totalCoeff < maxNumCoeff) {
        uint32_t w = br.show32();
        uint8_t t = (maxNumCoeff == 4) ? gTzDc[totalCoeff - 1][w >> 29] : gTz[totalCoeff - 1][w >> 23];
        if (!(t & 15)) return bad;
        if (!br.skip(t & 15)) return bad;
        zerosLeft = t >> 4;
    }
    int run[16];
    for (i = 0; i < totalCoeff - 1; i++) {
        if (zerosLeft > 0) {
            uint32_t w = br.show32();
            uint8_t t = gRun[(zerosLeft > 7 ? 7 : zerosLeft) - 1][w >> 21];
            if (!(t & 15)) return bad;
            int r = t >> 4;
            if (r > zerosLeft) return bad;
            if (!br.skip(t & 15)) return bad;
            run[i] = r;
            zerosLeft -= r;
        } else {
            run[i] = 0;
        }
    }
    // place levels: the last decoded level sits after `zerosLeft` leading zeros
    int pos = zerosLeft;
    uint32_t map = 0;
    for (i = totalCoeff - 1; i >= 0; i--) {
        if (i < totalCoeff - 1) pos += run[i] + 1;
        if (pos >= maxNumCoeff) return bad;
        int v = level[i];
        if (v > 32767 || v < -32768) return bad;
        out[pos] = (int16_t)v;
        map |= 1u << pos;
    }
    return CavlcResult{totalCoeff, map};
}

}  // namespace b200
