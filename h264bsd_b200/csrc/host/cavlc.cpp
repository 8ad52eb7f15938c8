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

// ---- expanded prefix LUTs.  entry: bits 0-4 length (0 = invalid), 5-6 trailingOnes, 7-11 totalCoeff.
// coeff_token in two levels of 8 bits (a flat 16-bit table is 128 KB per VLC table, and a short code followed by arbitrary
// bits lands anywhere in it): gTok1 by the first 8 bits holds the codes of up to 8 bits, or kTokLong | sub-table number;
// gTok2 by the next 8 bits holds the longer ones.
constexpr uint16_t kTokLong = 0x8000;
uint16_t gTok1[3][256];
uint16_t gTok2[3][16][256];
uint16_t gTokDc[1 << 8];
// entry: low nibble = length (0 invalid), high nibble = value
uint8_t gTz[15][1 << 9];
uint8_t gTzDc[3][1 << 3];
uint8_t gRun[7][1 << 11];  // value up to 14 -> (value<<4)|len  len up to 11 fits in 4 bits
std::once_flag gOnce;

void buildTables() {
    std::memset(gTok1, 0, sizeof gTok1);
    std::memset(gTok2, 0, sizeof gTok2);
    for (int t = 0; t < 3; t++) {
        int nSub = 0;
        for (int t1 = 0; t1 < 4; t1++)
            for (int tc = 0; tc < 17; tc++) {
                int len = kTokLen[t][t1][tc];
                if (!len) continue;
                uint32_t first = (uint32_t)kTokCode[t][t1][tc] << (16 - len);   // left-aligned in 16 bits
                uint16_t e = (uint16_t)(len | (t1 << 5) | (tc << 7));
                if (len <= 8) {
                    for (uint32_t k = 0; k < (1u << (8 - len)); k++) gTok1[t][(first >> 8) + k] = e;
                } else {
                    uint16_t &l1 = gTok1[t][first >> 8];
                    if (!(l1 & kTokLong)) l1 = (uint16_t)(kTokLong | nSub++);
                    uint16_t *sub = gTok2[t][l1 & 0xFF];
                    for (uint32_t k = 0; k < (1u << (16 - len)); k++) sub[(first & 0xFF) + k] = e;
                }
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

// coeff_token for VLC table t from the next 16 stream bits
static inline uint32_t tokLookup(int t, uint32_t prefix16) {
    uint32_t e = gTok1[t][prefix16 >> 8];
    if (e & kTokLong) e = gTok2[t][e & 0xFF][prefix16 & 0xFF];
    return e;
}

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
            return tokLookup(index < 2 ? 0 : index < 4 ? 1 : 2, prefix16);
        case 1:  // total_zeros 4x4, index = totalCoeff
            return gTz[index - 1][prefix16 >> 7];
        case 2:  // total_zeros chroma DC
            return gTzDc[index - 1][prefix16 >> 13];
        case 3:  // run_before, index = zerosLeft
            return gRun[(index > 7 ? 7 : index) - 1][prefix16 >> 5];
    }
    return 0;
}

CavlcResult cavlcResidualBlock(BitReader &stream, int16_t *out, int nC, int maxNumCoeff) {
    const CavlcResult bad{-1, 0, 0};
    // A 64-bit window of the stream, refilled when fewer than 32 of its bits are left.  Reading past the end of the NAL
    // yields zeros (h264bsdShowBits32) and is an error (END_OF_STREAM): checked once, at the end -- a block that crossed the
    // end is rejected whatever it decoded to, and an error ends the slice, so nothing of it is used.
    uint64_t pos = stream.pos();
    uint64_t win = stream.window(pos);
    unsigned used = 0;
    auto peek = [&]() -> uint32_t {
        if (used > 25) { pos += used; win = stream.window(pos); used = 0; }
        return (uint32_t)((win << used) >> 32);
    };

    uint32_t bits = peek();
    uint32_t e;
    if (nC < 0) {
        e = gTokDc[bits >> 24];
    } else if (nC < 8) {
        e = tokLookup(nC < 2 ? 0 : nC < 4 ? 1 : 2, bits >> 16);
    } else {
        uint32_t c = bits >> 26;
        if (c == 3) {
            e = 6;
        } else {
            uint32_t t1 = c & 3, tc = (c >> 2) + 1;
            e = (t1 > tc) ? 0 : (6 | (t1 << 5) | (tc << 7));
        }
    }
    const unsigned tokLen = e & 31;
    if (!tokLen) return bad;
    const int totalCoeff = (int)(e >> 7);
    const int trailingOnes = (int)((e >> 5) & 3);
    used += tokLen;
    if (totalCoeff > maxNumCoeff) return bad;
    if (totalCoeff == 0) {
        if (pos + used > stream.bitsTotal()) return bad;
        stream.seek(pos + used);
        return CavlcResult{0, 0, 0};
    }

    int level[16];
    int i = 0;
    if (trailingOnes) {
        const uint32_t signs = peek() >> (32 - trailingOnes);
        used += (unsigned)trailingOnes;
        for (int k = trailingOnes - 1; k >= 0; k--) level[i++] = (signs >> k) & 1 ? -1 : 1;
    }
    int suffixLength = (totalCoeff > 10 && trailingOnes < 3) ? 1 : 0;
    for (; i < totalCoeff; i++) {
        uint32_t w = peek();
        if ((w >> 16) == 0) return bad;  // level_prefix > 15 does not exist in Baseline
        const int prefix = __builtin_clz(w);
        used += (unsigned)prefix + 1;
        int levelCode = (prefix < 15 ? prefix : 15) << suffixLength;
        int suffixSize = suffixLength;
        if (prefix == 14 && suffixLength == 0) suffixSize = 4;
        if (prefix == 15) suffixSize = 12;
        if (suffixSize) {
            levelCode += (int)(peek() >> (32 - suffixSize));
            used += (unsigned)suffixSize;
        }
        if (prefix == 15 && suffixLength == 0) levelCode += 15;
        if (i == trailingOnes && trailingOnes < 3) levelCode += 2;
        const int mag = (levelCode + 2) >> 1;
        if (suffixLength == 0) suffixLength = 1;
        if (mag > (3 << (suffixLength - 1)) && suffixLength < 6) suffixLength++;
        level[i] = (levelCode & 1) ? -mag : mag;
    }

    int zerosLeft = 0;
    if (totalCoeff < maxNumCoeff) {
        const uint32_t w = peek();
        const uint8_t t = (maxNumCoeff == 4) ? gTzDc[totalCoeff - 1][w >> 29] : gTz[totalCoeff - 1][w >> 23];
        if (!(t & 15)) return bad;
        used += t & 15;
        zerosLeft = t >> 4;
    }
    // place levels from the highest frequency down: the first decoded level sits at position totalCoeff - 1 + total_zeros
    // (the tables keep total_zeros <= 16 - totalCoeff, so at <= 15.  For a 15-coefficient block that allows position 15,
    // one past its end: the reference does not check and the level lands in the first entry of the next block of its
    // level[26][16] array, h264bsd_cavlc.c:889-896 -- the caller's array has the same geometry, see parseResidual)
    int at = totalCoeff - 1 + zerosLeft;
    uint32_t map = 0, sumAbs = 0;
    for (i = 0; i < totalCoeff; i++) {
        const int v = level[i];
        if (v > 32767 || v < -32768) return bad;
        out[at] = (int16_t)v;
        map |= 1u << at;
        sumAbs += (uint32_t)(v < 0 ? -v : v);
        if (i == totalCoeff - 1) break;
        int r = 0;
        if (zerosLeft > 0) {
            const uint32_t w = peek();
            unsigned len;
            if (zerosLeft > 6) {
                // Table 9-10, last column: 111..001 -> 0..6, then one more leading zero per step
                if (w >> 29) { r = 7 - (int)(w >> 29); len = 3; }
                else { const int z = __builtin_clz(w | 1u); r = z + 4; len = (unsigned)z + 1; if (r > 14) return bad; }
            } else {
                const uint8_t t = gRun[zerosLeft - 1][w >> 21];
                if (!(t & 15)) return bad;
                r = t >> 4;
                len = t & 15;
            }
            if (r > zerosLeft) return bad;
            used += len;
            zerosLeft -= r;
        }
        at -= r + 1;
    }
    if (pos + used > stream.bitsTotal()) return bad;
    stream.seek(pos + used);
    return CavlcResult{totalCoeff, map, sumAbs};
}

}  // namespace b200
