// kdpt_scene.cpp — host-side mirror of the reference's scene / OBJ loading and KD-tree build (pure C++, no CUDA):
// what src/scene.cpp, src/KDtree.cpp, src/KDnode.cpp, src/objmesh.cpp + tiny_obj_loader v1.0.0 and
// src/utilities.cpp do between main() and pathtraceInit(). Output = exactly the arrays pathtraceInit reads
// (kdpt_scene), byte-identical to the reference loader's for the same files (tests/test_scene_loader_cpu.py).
//
// Written from the reference's behaviour, not its code: one pass, flat arrays, no per-triangle heap objects (the
// reference allocates a 116-byte KDN::Triangle per face and copies pointer vectors per node).
//
// Arithmetic that decides bytes is kept in the reference's order and precision (this file is compiled with
// -ffp-contract=off): glm 0.9.6.3 translate/rotate/scale/operator*/inverse/inverseTranspose, tinyobj's own decimal
// parser (tryParseDouble: it is NOT correctly rounded, so strtod would give different vertices), KDnode's
// float/double mix in the split predicate.
#include "../../include/kdpt.h"
#include "kdpt_bulk.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace kdpt {
void setError(const std::string& msg);
// kdpt_build.cu: the same build, level-synchronous on the GPU
int buildTreeDevice(const float* tri18, const int* mtl, int n, int maxdepth, int device, std::vector<kdpt_node_bare>& nodes,
                    kdpt::BulkVector<kdpt_tri_bare>& tris, kdpt::BulkVector<int>& triSource, float* deviceMs);
}

namespace {

const float kPI = 3.1415926535897932384626422832795028841971f;   // src/utilities.h:12

// ---- line reading / tokenising (utilities.cpp:78-139) -----------------------------------------------------------
bool safeGetline(std::istream& is, std::string& t)
{
    t.clear();
    std::streambuf* sb = is.rdbuf();
    for (;;) {
        int c = sb->sbumpc();
        switch (c) {
        case '\n': return true;
        case '\r': if (sb->sgetc() == '\n') sb->sbumpc(); return true;
        case EOF: if (t.empty()) is.setstate(std::ios::eofbit); return !t.empty();
        default: t += (char)c;
        }
    }
}
std::vector<std::string> tokenize(const std::string& s)
{
    std::istringstream ss(s);
    std::vector<std::string> r; std::string w;
    while (ss >> w) r.push_back(w);
    return r;
}
float tokf(const std::vector<std::string>& t, size_t i) { return i < t.size() ? (float)atof(t[i].c_str()) : 0.0f; }

// ---- glm 0.9.6.3 mat4 (column-major float[16]) --------------------------------------------------------------------
struct M4 { float c[4][4]; };   // c[col][row]
M4 identity() { M4 m; memset(&m, 0, sizeof m); for (int i = 0; i < 4; i++) m.c[i][i] = 1.0f; return m; }
// gtc/matrix_transform.inl translate: Result[3] = m[0]*v[0] + m[1]*v[1] + m[2]*v[2] + m[3]
M4 translate(const M4& m, const float* v)
{
    M4 r = m;
    for (int k = 0; k < 4; k++) r.c[3][k] = ((m.c[0][k] * v[0] + m.c[1][k] * v[1]) + m.c[2][k] * v[2]) + m.c[3][k];
    return r;
}
// gtc/matrix_transform.inl rotate
M4 rotate(const M4& m, float angle, const float* v)
{
    const float a = angle, c = cosf(a), s = sinf(a);
    float inv = 1.0f / sqrtf((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]);
    float axis[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
    float temp[3] = {(1.0f - c) * axis[0], (1.0f - c) * axis[1], (1.0f - c) * axis[2]};
    float R[3][3];
    R[0][0] = c + temp[0] * axis[0];
    R[0][1] = 0 + temp[0] * axis[1] + s * axis[2];
    R[0][2] = 0 + temp[0] * axis[2] - s * axis[1];
    R[1][0] = 0 + temp[1] * axis[0] - s * axis[2];
    R[1][1] = c + temp[1] * axis[1];
    R[1][2] = 0 + temp[1] * axis[2] + s * axis[0];
    R[2][0] = 0 + temp[2] * axis[0] + s * axis[1];
    R[2][1] = 0 + temp[2] * axis[1] - s * axis[0];
    R[2][2] = c + temp[2] * axis[2];
    M4 r;
    for (int j = 0; j < 3; j++)
        for (int k = 0; k < 4; k++) r.c[j][k] = (m.c[0][k] * R[j][0] + m.c[1][k] * R[j][1]) + m.c[2][k] * R[j][2];
    for (int k = 0; k < 4; k++) r.c[3][k] = m.c[3][k];
    return r;
}
M4 scale(const M4& m, const float* v)
{
    M4 r;
    for (int j = 0; j < 3; j++) for (int k = 0; k < 4; k++) r.c[j][k] = m.c[j][k] * v[j];
    for (int k = 0; k < 4; k++) r.c[3][k] = m.c[3][k];
    return r;
}
// detail/type_mat4x4.inl:686-704
M4 mul(const M4& a, const M4& b)
{
    M4 r;
    for (int j = 0; j < 4; j++)
        for (int k = 0; k < 4; k++)
            r.c[j][k] = ((a.c[0][k] * b.c[j][0] + a.c[1][k] * b.c[j][1]) + a.c[2][k] * b.c[j][2]) + a.c[3][k] * b.c[j][3];
    return r;
}
// detail/type_mat4x4.inl compute_inverse
M4 inverse(const M4& M)
{
    auto m = [&](int a, int b) { return M.c[a][b]; };
    float Coef00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3);
    float Coef02 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3);
    float Coef03 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
    float Coef04 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3);
    float Coef06 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3);
    float Coef07 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3);
    float Coef08 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2);
    float Coef10 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2);
    float Coef11 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
    float Coef12 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3);
    float Coef14 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3);
    float Coef15 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3);
    float Coef16 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2);
    float Coef18 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2);
    float Coef19 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
    float Coef20 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1);
    float Coef22 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1);
    float Coef23 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
    float Fac0[4] = {Coef00, Coef00, Coef02, Coef03}, Fac1[4] = {Coef04, Coef04, Coef06, Coef07};
    float Fac2[4] = {Coef08, Coef08, Coef10, Coef11}, Fac3[4] = {Coef12, Coef12, Coef14, Coef15};
    float Fac4[4] = {Coef16, Coef16, Coef18, Coef19}, Fac5[4] = {Coef20, Coef20, Coef22, Coef23};
    float Vec0[4] = {m(1, 0), m(0, 0), m(0, 0), m(0, 0)}, Vec1[4] = {m(1, 1), m(0, 1), m(0, 1), m(0, 1)};
    float Vec2[4] = {m(1, 2), m(0, 2), m(0, 2), m(0, 2)}, Vec3[4] = {m(1, 3), m(0, 3), m(0, 3), m(0, 3)};
    const float SignA[4] = {+1, -1, +1, -1}, SignB[4] = {-1, +1, -1, +1};
    M4 I;
    for (int k = 0; k < 4; k++) {
        float Inv0 = (Vec1[k] * Fac0[k] - Vec2[k] * Fac1[k]) + Vec3[k] * Fac2[k];
        float Inv1 = (Vec0[k] * Fac0[k] - Vec2[k] * Fac3[k]) + Vec3[k] * Fac4[k];
        float Inv2 = (Vec0[k] * Fac1[k] - Vec1[k] * Fac3[k]) + Vec3[k] * Fac5[k];
        float Inv3 = (Vec0[k] * Fac2[k] - Vec1[k] * Fac4[k]) + Vec2[k] * Fac5[k];
        I.c[0][k] = Inv0 * SignA[k]; I.c[1][k] = Inv1 * SignB[k]; I.c[2][k] = Inv2 * SignA[k]; I.c[3][k] = Inv3 * SignB[k];
    }
    float Dot0[4] = {M.c[0][0] * I.c[0][0], M.c[0][1] * I.c[1][0], M.c[0][2] * I.c[2][0], M.c[0][3] * I.c[3][0]};
    float Dot1 = (Dot0[0] + Dot0[1]) + (Dot0[2] + Dot0[3]);
    float OneOverDeterminant = 1.0f / Dot1;
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) I.c[j][k] = I.c[j][k] * OneOverDeterminant;
    return I;
}
// gtc/matrix_inverse.inl:95-146
M4 inverseTranspose(const M4& M)
{
    auto m = [&](int a, int b) { return M.c[a][b]; };
    float S00 = m(2, 2) * m(3, 3) - m(3, 2) * m(2, 3), S01 = m(2, 1) * m(3, 3) - m(3, 1) * m(2, 3);
    float S02 = m(2, 1) * m(3, 2) - m(3, 1) * m(2, 2), S03 = m(2, 0) * m(3, 3) - m(3, 0) * m(2, 3);
    float S04 = m(2, 0) * m(3, 2) - m(3, 0) * m(2, 2), S05 = m(2, 0) * m(3, 1) - m(3, 0) * m(2, 1);
    float S06 = m(1, 2) * m(3, 3) - m(3, 2) * m(1, 3), S07 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3);
    float S08 = m(1, 1) * m(3, 2) - m(3, 1) * m(1, 2), S09 = m(1, 0) * m(3, 3) - m(3, 0) * m(1, 3);
    float S10 = m(1, 0) * m(3, 2) - m(3, 0) * m(1, 2), S11 = m(1, 1) * m(3, 3) - m(3, 1) * m(1, 3);
    float S12 = m(1, 0) * m(3, 1) - m(3, 0) * m(1, 1), S13 = m(1, 2) * m(2, 3) - m(2, 2) * m(1, 3);
    float S14 = m(1, 1) * m(2, 3) - m(2, 1) * m(1, 3), S15 = m(1, 1) * m(2, 2) - m(2, 1) * m(1, 2);
    float S16 = m(1, 0) * m(2, 3) - m(2, 0) * m(1, 3), S17 = m(1, 0) * m(2, 2) - m(2, 0) * m(1, 2);
    float S18 = m(1, 0) * m(2, 1) - m(2, 0) * m(1, 1);
    M4 I;
    I.c[0][0] = +((m(1, 1) * S00 - m(1, 2) * S01) + m(1, 3) * S02);
    I.c[0][1] = -((m(1, 0) * S00 - m(1, 2) * S03) + m(1, 3) * S04);
    I.c[0][2] = +((m(1, 0) * S01 - m(1, 1) * S03) + m(1, 3) * S05);
    I.c[0][3] = -((m(1, 0) * S02 - m(1, 1) * S04) + m(1, 2) * S05);
    I.c[1][0] = -((m(0, 1) * S00 - m(0, 2) * S01) + m(0, 3) * S02);
    I.c[1][1] = +((m(0, 0) * S00 - m(0, 2) * S03) + m(0, 3) * S04);
    I.c[1][2] = -((m(0, 0) * S01 - m(0, 1) * S03) + m(0, 3) * S05);
    I.c[1][3] = +((m(0, 0) * S02 - m(0, 1) * S04) + m(0, 2) * S05);
    I.c[2][0] = +((m(0, 1) * S06 - m(0, 2) * S07) + m(0, 3) * S08);
    I.c[2][1] = -((m(0, 0) * S06 - m(0, 2) * S09) + m(0, 3) * S10);
    I.c[2][2] = +((m(0, 0) * S11 - m(0, 1) * S09) + m(0, 3) * S12);
    I.c[2][3] = -((m(0, 0) * S08 - m(0, 1) * S10) + m(0, 2) * S12);
    I.c[3][0] = -((m(0, 1) * S13 - m(0, 2) * S14) + m(0, 3) * S15);
    I.c[3][1] = +((m(0, 0) * S13 - m(0, 2) * S16) + m(0, 3) * S17);
    I.c[3][2] = -((m(0, 0) * S14 - m(0, 1) * S16) + m(0, 3) * S18);
    I.c[3][3] = +((m(0, 0) * S15 - m(0, 1) * S17) + m(0, 2) * S18);
    float Determinant = ((+m(0, 0) * I.c[0][0] + m(0, 1) * I.c[0][1]) + m(0, 2) * I.c[0][2]) + m(0, 3) * I.c[0][3];
    for (int j = 0; j < 4; j++) for (int k = 0; k < 4; k++) I.c[j][k] = I.c[j][k] / Determinant;
    return I;
}
// utilities.cpp:65-72
M4 buildTransformationMatrix(const float* t, const float* r, const float* s)
{
    M4 T = translate(identity(), t);
    const float ax[3] = {1, 0, 0}, ay[3] = {0, 1, 0}, az[3] = {0, 0, 1};
    M4 R = rotate(identity(), r[0] * kPI / 180, ax);
    R = mul(R, rotate(identity(), r[1] * kPI / 180, ay));
    R = mul(R, rotate(identity(), r[2] * kPI / 180, az));
    M4 S = scale(identity(), s);
    return mul(mul(T, R), S);
}

// ---- tinyobjloader v1.0.0 number parsing (tiny_obj_loader.cpp:158-281) --------------------------------------------
bool isDigit(char x) { return (unsigned)(x - '0') < 10u; }
bool tryParseDouble(const char* s, const char* s_end, double* result)
{
    if (s >= s_end) return false;
    double mantissa = 0.0; int exponent = 0; char sign = '+', exp_sign = '+';
    const char* curr = s; int read = 0; bool end_not_reached = false;
    if (*curr == '+' || *curr == '-') { sign = *curr; curr++; }
    else if (!isDigit(*curr)) return false;
    end_not_reached = (curr != s_end);
    while (end_not_reached && isDigit(*curr)) { mantissa *= 10; mantissa += (int)(*curr - 0x30); curr++; read++; end_not_reached = (curr != s_end); }
    if (read == 0) return false;
    if (end_not_reached) {
        bool parseExp = false;
        if (*curr == '.') {
            curr++; read = 1; end_not_reached = (curr != s_end);
            while (end_not_reached && isDigit(*curr)) {
                mantissa += (int)(*curr - 0x30) * pow(10.0, -read);
                read++; curr++; end_not_reached = (curr != s_end);
            }
            parseExp = end_not_reached;
        } else if (*curr == 'e' || *curr == 'E') parseExp = true;
        if (parseExp && (*curr == 'e' || *curr == 'E')) {
            curr++;
            end_not_reached = (curr != s_end);
            if (end_not_reached && (*curr == '+' || *curr == '-')) { exp_sign = *curr; curr++; }
            else if (!isDigit(*curr)) return false;
            read = 0; end_not_reached = (curr != s_end);
            while (end_not_reached && isDigit(*curr)) { exponent *= 10; exponent += (int)(*curr - 0x30); curr++; read++; end_not_reached = (curr != s_end); }
            exponent *= (exp_sign == '+' ? 1 : -1);
            if (read == 0) return false;
        }
    }
    *result = (sign == '+' ? 1 : -1) * ldexp(mantissa * pow(5.0, exponent), exponent);
    return true;
}
float parseFloat(const char** token, double def = 0.0)
{
    (*token) += strspn(*token, " \t");
    const char* end = (*token) + strcspn(*token, " \t\r");
    double val = def;
    tryParseDouble(*token, end, &val);
    *token = end;
    return (float)val;
}
int fixIndex(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }

struct ObjMaterial { float ambient[3] = {0, 0, 0}, diffuse[3] = {0, 0, 0}, specular[3] = {0, 0, 0}, transmittance[3] = {0, 0, 0}; int illum = 0; float ior = 1.f; std::string name; };
struct ObjData {
    std::vector<float> v, vn;
    std::vector<std::vector<int>> shapes;     // per shape: vertex indices, 3 per triangle (fan-triangulated)
    std::vector<ObjMaterial> materials;
};

void trimLine(std::string& l)
{
    if (!l.empty() && l[l.size() - 1] == '\n') l.erase(l.size() - 1);
    if (!l.empty() && l[l.size() - 1] == '\r') l.erase(l.size() - 1);
}

// LoadMtl (tiny_obj_loader.cpp:486-797), only the fields the reference consumes
void loadMtl(const std::string& path, ObjData& out, std::map<std::string, int>& matMap)
{
    std::ifstream in(path.c_str());
    ObjMaterial cur;
    while (in.peek() != -1) {
        std::string l;
        safeGetline(in, l);
        if (!l.empty()) l = l.substr(0, l.find_last_not_of(" \t") + 1);
        trimLine(l);
        if (l.empty()) continue;
        const char* t = l.c_str();
        t += strspn(t, " \t");
        if (t[0] == '\0' || t[0] == '#') continue;
        auto sp = [&](int k) { return t[k] == ' ' || t[k] == '\t'; };
        if (strncmp(t, "newmtl", 6) == 0 && sp(6)) {
            if (!cur.name.empty()) { matMap.insert(std::make_pair(cur.name, (int)out.materials.size())); out.materials.push_back(cur); }
            cur = ObjMaterial();
            char namebuf[4096]; namebuf[0] = 0;
            sscanf(t + 7, "%4095s", namebuf);
            cur.name = namebuf;
            continue;
        }
        auto f3 = [&](float* dst) { const char* p = t + 2; dst[0] = parseFloat(&p); dst[1] = parseFloat(&p); dst[2] = parseFloat(&p); };
        if (t[0] == 'K' && t[1] == 'a' && sp(2)) { f3(cur.ambient); continue; }
        if (t[0] == 'K' && t[1] == 'd' && sp(2)) { f3(cur.diffuse); continue; }
        if (t[0] == 'K' && t[1] == 's' && sp(2)) { f3(cur.specular); continue; }
        if ((t[0] == 'K' && t[1] == 't' && sp(2)) || (t[0] == 'T' && t[1] == 'f' && sp(2))) { f3(cur.transmittance); continue; }
        if (t[0] == 'N' && t[1] == 'i' && sp(2)) { const char* p = t + 2; cur.ior = parseFloat(&p); continue; }
        if (strncmp(t, "illum", 5) == 0 && sp(5)) { const char* p = t + 6; p += strspn(p, " \t"); cur.illum = atoi(p); continue; }
    }
    // the last material is flushed unconditionally (also the nameless default one of a missing file)
    matMap.insert(std::make_pair(cur.name, (int)out.materials.size()));
    out.materials.push_back(cur);
}

// LoadObj (tiny_obj_loader.cpp:842-1156), triangulate = true; shapes split at g / o lines only
bool loadObj(const std::string& path, const std::string& mtlBase, ObjData& out, std::string& err)
{
    std::ifstream in(path.c_str());
    if (!in) { err = "Cannot open file [" + path + "]"; return false; }
    std::vector<std::vector<int>> faceGroup;
    std::vector<int> shape;
    std::map<std::string, int> matMap;
    int material = -1;
    auto exportGroup = [&]() -> bool {
        if (faceGroup.empty()) return false;
        for (auto& face : faceGroup) {
            if (face.size() < 2) continue;          // the reference reads face[1] unconditionally; malformed input
            int i0 = face[0], i2 = face[1];
            for (size_t k = 2; k < face.size(); k++) { int i1 = i2; i2 = face[k]; shape.push_back(i0); shape.push_back(i1); shape.push_back(i2); }
        }
        return true;
    };
    while (in.peek() != -1) {
        std::string l;
        safeGetline(in, l);
        trimLine(l);
        if (l.empty()) continue;
        const char* t = l.c_str();
        t += strspn(t, " \t");
        if (t[0] == '\0' || t[0] == '#') continue;
        auto sp = [&](int k) { return t[k] == ' ' || t[k] == '\t'; };
        if (t[0] == 'v' && sp(1)) { const char* p = t + 2; for (int k = 0; k < 3; k++) out.v.push_back(parseFloat(&p)); continue; }
        if (t[0] == 'v' && t[1] == 'n' && sp(2)) { const char* p = t + 3; for (int k = 0; k < 3; k++) out.vn.push_back(parseFloat(&p)); continue; }
        if (t[0] == 'v' && t[1] == 't' && sp(2)) continue;
        if (t[0] == 'f' && sp(1)) {
            const char* p = t + 2;
            p += strspn(p, " \t");
            std::vector<int> face;
            while (!(p[0] == '\r' || p[0] == '\n' || p[0] == '\0')) {
                face.push_back(fixIndex(atoi(p), (int)(out.v.size() / 3)));     // only vertex_index is kept (scene.cpp:647)
                p += strcspn(p, " \t\r");
                p += strspn(p, " \t\r");
            }
            faceGroup.push_back(face);
            continue;
        }
        if (strncmp(t, "usemtl", 6) == 0 && sp(6)) {
            char namebuf[4096]; namebuf[0] = 0;
            sscanf(t + 7, "%4095s", namebuf);
            int id = -1;
            auto it = matMap.find(namebuf);
            if (it != matMap.end()) id = it->second;
            if (id != material) { exportGroup(); faceGroup.clear(); material = id; }
            continue;
        }
        if (strncmp(t, "mtllib", 6) == 0 && sp(6)) {
            char namebuf[4096]; namebuf[0] = 0;
            sscanf(t + 7, "%4095s", namebuf);
            loadMtl(mtlBase.empty() ? std::string(namebuf) : mtlBase + namebuf, out, matMap);
            continue;
        }
        if ((t[0] == 'g' || t[0] == 'o') && sp(1)) {
            if (exportGroup()) out.shapes.push_back(shape);
            shape.clear(); faceGroup.clear();
            continue;
        }
    }
    if (exportGroup()) out.shapes.push_back(shape);
    return true;
}

// ---- KD build + flatten (KDnode.cpp:112-249, scene.cpp:866-968) ---------------------------------------------------
struct Builder {
    const float* tri18; const int* mtl; int maxdepth;
    std::vector<float> tmin, tmax;                  // per-triangle bounds (KDnode.h:216-225)
    std::vector<kdpt_node_bare> nodes;
    std::vector<int> leafTri;

    int newNode()
    {
        kdpt_node_bare n; memset(&n, 0, sizeof n);
        n.ID = (int)nodes.size(); n.parentID = n.leftID = n.rightID = -1; n.triIdStart = n.triIdSize = -1;
        nodes.push_back(n);
        return n.ID;
    }
    void leaf(int id, const int* tris, int num)
    {
        if (num <= 0) return;
        nodes[id].triIdStart = (int)leafTri.size();
        nodes[id].triIdSize = num;
        leafTri.insert(leafTri.end(), tris, tris + num);
    }
    // `center`: the root keeps the centre computed BEFORE the +-0.001 pad (updateBbox never refreshes it)
    void split(int id, std::vector<int>& tris, int depth, const float* center)
    {
        const int num = (int)tris.size();
        if (num <= 2 || depth > maxdepth) { leaf(id, tris.data(), num); return; }
        const int level = depth % 3;
        const double c = (double)center[level];
        std::vector<int> L, R;
        if (num < (1 << 17)) {
            L.reserve(num / 2 + 16); R.reserve(num / 2 + 16);
            for (int i = 0; i < num; i++) {
                int t = tris[i];
                if ((double)tmin[3 * (size_t)t + level] < c + 0.0001) L.push_back(t);
                if ((double)tmax[3 * (size_t)t + level] >= c - 0.0001) R.push_back(t);
            }
        } else {
            // big nodes (the top of a multi-million-triangle tree): the classification is a per-triangle predicate, so
            // it is split over the host cores; pieces are concatenated in order, which keeps the lists (and therefore
            // every leaf's triangle order and every node ID) identical to the sequential build
            unsigned hw = std::thread::hardware_concurrency();
            const int nt = (int)(hw == 0 ? 4 : (hw > 32 ? 32 : hw));
            std::vector<std::vector<int>> Ls(nt), Rs(nt);
            std::vector<std::thread> th;
            for (int k = 0; k < nt; k++)
                th.emplace_back([&, k]() {
                    const int lo = (int)((long long)num * k / nt), hi = (int)((long long)num * (k + 1) / nt);
                    Ls[k].reserve((hi - lo) / 2 + 16); Rs[k].reserve((hi - lo) / 2 + 16);
                    for (int i = lo; i < hi; i++) {
                        int t = tris[i];
                        if ((double)tmin[3 * (size_t)t + level] < c + 0.0001) Ls[k].push_back(t);
                        if ((double)tmax[3 * (size_t)t + level] >= c - 0.0001) Rs[k].push_back(t);
                    }
                });
            for (auto& t : th) t.join();
            size_t nl = 0, nr = 0;
            for (int k = 0; k < nt; k++) { nl += Ls[k].size(); nr += Rs[k].size(); }
            L.reserve(nl); R.reserve(nr);
            for (int k = 0; k < nt; k++) { L.insert(L.end(), Ls[k].begin(), Ls[k].end()); R.insert(R.end(), Rs[k].begin(), Rs[k].end()); }
        }
        if ((int)L.size() == num || (int)R.size() == num) { leaf(id, tris.data(), num); return; }
        std::vector<int>().swap(tris);              // interior nodes drop their list (KDnode.cpp:246-247)
        float pmin[3], pmax[3];
        memcpy(pmin, nodes[id].mins, sizeof pmin); memcpy(pmax, nodes[id].maxs, sizeof pmax);
        const float cl = center[level];
        for (int side = 0; side < 2; side++) {
            std::vector<int>& S = side == 0 ? L : R;
            if (S.empty()) continue;
            int ch = newNode();
            kdpt_node_bare& n = nodes[ch];
            n.axis = (level + 1) % 3; n.parentID = id;
            memcpy(n.mins, pmin, sizeof pmin); memcpy(n.maxs, pmax, sizeof pmax);
            if (side == 0) { nodes[id].leftID = ch; n.maxs[level] = cl; n.splitPos = pmax[level]; }
            else { nodes[id].rightID = ch; n.mins[level] = cl; n.splitPos = pmin[level]; }
            float cc[3];
            for (int a = 0; a < 3; a++) cc[a] = (float)((double)(nodes[ch].mins[a] + nodes[ch].maxs[a]) / 2.0);
            split(ch, S, depth + 1, cc);
        }
    }
};

struct HostScene {
    std::vector<kdpt_node_bare> nodes;
    kdpt::BulkVector<kdpt_tri_bare> tris;
    kdpt::BulkVector<int> triSource;
    std::vector<kdpt_geom> geoms;
    std::vector<kdpt_material> materials;
    std::vector<int> offsets;
    kdpt_camera camera;
    int traceDepth = 0, iterations = 0;
    std::string imageName;
};

void buildTree(const float* tri18, const int* mtl, int n, int maxdepth, HostScene& hs)
{
    Builder b; b.tri18 = tri18; b.mtl = mtl; b.maxdepth = maxdepth;
    b.tmin.resize(3 * (size_t)n); b.tmax.resize(3 * (size_t)n);
    for (int i = 0; i < n; i++) {
        const float* t = tri18 + 18 * (size_t)i;
        for (int a = 0; a < 3; a++) {
            float c1 = t[a], c2 = t[3 + a], c3 = t[6 + a];
            b.tmin[3 * (size_t)i + a] = c1 < c2 ? (c1 < c3 ? c1 : c3) : (c2 < c3 ? c2 : c3);
            b.tmax[3 * (size_t)i + a] = c1 > c2 ? (c1 > c3 ? c1 : c3) : (c2 > c3 ? c2 : c3);
        }
    }
    int root = b.newNode();
    float mn[3] = {0, 0, 0}, mx[3] = {0, 0, 0}, center[3] = {0, 0, 0};
    if (n > 0) {
        for (int a = 0; a < 3; a++) { mn[a] = b.tmin[a]; mx[a] = b.tmax[a]; }
        for (int i = 1; i < n; i++)
            for (int a = 0; a < 3; a++) {
                float lo = b.tmin[3 * (size_t)i + a], hi = b.tmax[3 * (size_t)i + a];
                mn[a] = mn[a] > lo ? lo : mn[a];
                mx[a] = mx[a] < hi ? hi : mx[a];
            }
        for (int a = 0; a < 3; a++) center[a] = (float)((double)(mn[a] + mx[a]) / 2.0);
    }
    const float pad = 0.001f;
    for (int a = 0; a < 3; a++) { mn[a] -= pad; mx[a] += pad; b.nodes[root].mins[a] = mn[a]; b.nodes[root].maxs[a] = mx[a]; }
    std::vector<int> all((size_t)n);
    for (int i = 0; i < n; i++) all[i] = i;
    b.split(root, all, 0, center);
    hs.nodes.swap(b.nodes);
    hs.triSource.assign(b.leafTri.begin(), b.leafTri.end());
    hs.tris.resize(hs.triSource.size());
    // cacheTrianglesBare (scene.cpp:934-968): one 76-byte component-major record per leaf entry; entries are independent,
    // so big trees are filled by all host cores (which then also take the page faults of the ~76 B x entries array)
    auto fill = [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; i++) {
            const float* s = tri18 + 18 * (size_t)hs.triSource[i];
            kdpt_tri_bare& t = hs.tris[i];
            t.x1 = s[0]; t.y1 = s[1]; t.z1 = s[2]; t.x2 = s[3]; t.y2 = s[4]; t.z2 = s[5]; t.x3 = s[6]; t.y3 = s[7]; t.z3 = s[8];
            t.nx1 = s[9]; t.ny1 = s[10]; t.nz1 = s[11]; t.nx2 = s[12]; t.ny2 = s[13]; t.nz2 = s[14];
            t.nx3 = s[15]; t.ny3 = s[16]; t.nz3 = s[17];
            t.mtlIdx = mtl ? mtl[hs.triSource[i]] : 0;
        }
    };
    const size_t total = hs.triSource.size();
    unsigned hw = std::thread::hardware_concurrency();
    const int nt = total < (1u << 17) ? 1 : (int)(hw == 0 ? 4 : (hw > 32 ? 32 : hw));
    if (nt == 1) fill(0, total);
    else {
        std::vector<std::thread> th;
        for (int k = 0; k < nt; k++) th.emplace_back(fill, total * k / nt, total * (k + 1) / nt);
        for (auto& t : th) t.join();
    }
}

// ---- scene .txt (scene.cpp:7-61, 118-271) ---------------------------------------------------------------------------
bool loadSceneTxt(const std::string& path, HostScene& hs, std::string& err)
{
    std::ifstream in(path.c_str());
    if (!in.is_open()) { err = "cannot open scene file " + path; return false; }
    memset(&hs.camera, 0, sizeof hs.camera);
    while (in.good()) {
        std::string line;
        safeGetline(in, line);
        if (line.empty()) continue;
        std::vector<std::string> tok = tokenize(line);
        if (tok.empty()) continue;
        if (tok[0] == "MATERIAL") {
            if (tok.size() < 2 || atoi(tok[1].c_str()) != (int)hs.materials.size()) { err = "MATERIAL ID does not match expected number of materials"; return false; }
            kdpt_material m; memset(&m, 0, sizeof m);      // the reference leaves transmittance uninitialised (Appendix A #8)
            for (int i = 0; i < 7; i++) {
                safeGetline(in, line);
                std::vector<std::string> t = tokenize(line);
                if (t.empty()) continue;
                if (t[0] == "RGB") { m.color[0] = tokf(t, 1); m.color[1] = tokf(t, 2); m.color[2] = tokf(t, 3); }
                else if (t[0] == "SPECEX") m.specular_exponent = tokf(t, 1);
                else if (t[0] == "SPECRGB") { m.specular_color[0] = tokf(t, 1); m.specular_color[1] = tokf(t, 2); m.specular_color[2] = tokf(t, 3); }
                else if (t[0] == "REFL") m.hasReflective = tokf(t, 1);
                else if (t[0] == "REFR") m.hasRefractive = tokf(t, 1);
                else if (t[0] == "REFRIOR") m.indexOfRefraction = tokf(t, 1);
                else if (t[0] == "EMITTANCE") m.emittance = tokf(t, 1);
            }
            hs.materials.push_back(m);
        } else if (tok[0] == "OBJECT") {
            if (tok.size() < 2 || atoi(tok[1].c_str()) != (int)hs.geoms.size()) { err = "OBJECT ID does not match expected number of geoms"; return false; }
            kdpt_geom g; memset(&g, 0, sizeof g);
            safeGetline(in, line);
            if (!line.empty() && in.good()) {
                if (line == "sphere") g.type = 0;
                else if (line == "cube") g.type = 1;
            }
            safeGetline(in, line);
            if (!line.empty() && in.good()) { std::vector<std::string> t = tokenize(line); if (t.size() > 1) g.materialid = atoi(t[1].c_str()); }
            safeGetline(in, line);
            while (!line.empty() && in.good()) {
                std::vector<std::string> t = tokenize(line);
                if (!t.empty()) {
                    float* dst = t[0] == "TRANS" ? g.translation : t[0] == "ROTAT" ? g.rotation : t[0] == "SCALE" ? g.scale : nullptr;
                    if (dst) { dst[0] = tokf(t, 1); dst[1] = tokf(t, 2); dst[2] = tokf(t, 3); }
                }
                safeGetline(in, line);
            }
            M4 T = buildTransformationMatrix(g.translation, g.rotation, g.scale);
            M4 I = inverse(T), IT = inverseTranspose(T);
            memcpy(g.transform, T.c, sizeof g.transform); memcpy(g.inverseTransform, I.c, sizeof g.inverseTransform);
            memcpy(g.invTranspose, IT.c, sizeof g.invTranspose);
            hs.geoms.push_back(g);
        } else if (tok[0] == "CAMERA") {
            kdpt_camera& c = hs.camera;
            float fovy = 0.f;
            for (int i = 0; i < 5; i++) {
                safeGetline(in, line);
                std::vector<std::string> t = tokenize(line);
                if (t.empty()) continue;
                if (t[0] == "RES") { c.resolution[0] = t.size() > 1 ? atoi(t[1].c_str()) : 0; c.resolution[1] = t.size() > 2 ? atoi(t[2].c_str()) : 0; }
                else if (t[0] == "FOVY") fovy = tokf(t, 1);
                else if (t[0] == "ITERATIONS") hs.iterations = t.size() > 1 ? atoi(t[1].c_str()) : 0;
                else if (t[0] == "DEPTH") hs.traceDepth = t.size() > 1 ? atoi(t[1].c_str()) : 0;
                else if (t[0] == "FILE") hs.imageName = t.size() > 1 ? t[1] : "";
            }
            safeGetline(in, line);
            while (!line.empty() && in.good()) {
                std::vector<std::string> t = tokenize(line);
                if (!t.empty()) {
                    float* dst = t[0] == "EYE" ? c.position : t[0] == "LOOKAT" ? c.lookAt : t[0] == "UP" ? c.up : nullptr;
                    if (dst) { dst[0] = tokf(t, 1); dst[1] = tokf(t, 2); dst[2] = tokf(t, 3); }
                }
                safeGetline(in, line);
            }
            float yscaled = tanf(fovy * (kPI / 180));
            float xscaled = (yscaled * c.resolution[0]) / c.resolution[1];
            float fovx = (atanf(xscaled) * 180) / kPI;
            c.fov[0] = fovx; c.fov[1] = fovy;
            // scene.cpp:221 computes `right` from the still zero `view` (Appendix A #9): cross(0, up) normalised = NaN;
            // it is recomputed by kdpt_camera_first_frame like the reference's runCuda() does
            {
                float v[3] = {c.view[0], c.view[1], c.view[2]}, u[3] = {c.up[0], c.up[1], c.up[2]};
                float r[3] = {v[1] * u[2] - u[1] * v[2], v[2] * u[0] - u[2] * v[0], v[0] * u[1] - u[0] * v[1]};
                float s = 1.0f / sqrtf((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2]);
                c.right[0] = r[0] * s; c.right[1] = r[1] * s; c.right[2] = r[2] * s;
            }
            c.pixelLength[0] = 2 * xscaled / (float)c.resolution[0];
            c.pixelLength[1] = 2 * yscaled / (float)c.resolution[1];
            float d[3] = {c.lookAt[0] - c.position[0], c.lookAt[1] - c.position[1], c.lookAt[2] - c.position[2]};
            float s = 1.0f / sqrtf((d[0] * d[0] + d[1] * d[1]) + d[2] * d[2]);
            c.view[0] = d[0] * s; c.view[1] = d[1] * s; c.view[2] = d[2] * s;
        }
    }
    return true;
}

// scene.cpp:579-901: OBJ -> per-shape Material, triangles (normals fetched by VERTEX index), KD tree, flat arrays
bool attachObj(const std::string& objPath, const std::string& mtlDir, int kdDepth, int synthNormals, HostScene& hs, std::string& err)
{
    ObjData obj;
    if (!loadObj(objPath, mtlDir, obj, err)) return false;
    const size_t nv = obj.v.size() / 3;
    std::vector<float> normals;
    if (obj.vn.size() >= obj.v.size()) normals = obj.vn;
    else if (synthNormals) {
        // `vn`-less meshes (scenes/dragon.obj): the reference would read out of bounds. Area-weighted vertex normals.
        std::vector<double> acc(3 * nv, 0.0);
        for (auto& sh : obj.shapes)
            for (size_t j = 0; j + 2 < sh.size(); j += 3) {
                int a = sh[j], b = sh[j + 1], c = sh[j + 2];
                if ((size_t)a >= nv || (size_t)b >= nv || (size_t)c >= nv || a < 0 || b < 0 || c < 0) continue;
                double e1[3], e2[3];
                for (int k = 0; k < 3; k++) { e1[k] = (double)obj.v[3 * b + k] - obj.v[3 * a + k]; e2[k] = (double)obj.v[3 * c + k] - obj.v[3 * a + k]; }
                double n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
                for (int k = 0; k < 3; k++) { acc[3 * a + k] += n[k]; acc[3 * b + k] += n[k]; acc[3 * c + k] += n[k]; }
            }
        normals.resize(3 * nv);
        for (size_t i = 0; i < nv; i++) {
            double l = sqrt(acc[3 * i] * acc[3 * i] + acc[3 * i + 1] * acc[3 * i + 1] + acc[3 * i + 2] * acc[3 * i + 2]);
            for (int k = 0; k < 3; k++) normals[3 * i + k] = l > 0 ? (float)(acc[3 * i + k] / l) : (k == 1 ? 1.f : 0.f);
        }
    } else {
        err = "OBJ has fewer normals than vertices: the reference fetches normals by vertex index (scene.cpp:559-567); pass synth_normals to generate them";
        return false;
    }
    const int numShapes = (int)obj.shapes.size();
    for (int i = 0; i < numShapes; i++) {
        kdpt_material m; memset(&m, 0, sizeof m);
        if ((int)obj.materials.size() > i) {
            const ObjMaterial& om = obj.materials[i];        // matched to shapes BY POSITION (scene.cpp:718-720)
            for (int k = 0; k < 3; k++) m.color[k] = om.ambient[k] > om.diffuse[k] ? om.ambient[k] : om.diffuse[k];
            if (om.illum <= 2) { /* diffuse */ }
            else if (om.illum == 3) { m.specular_exponent = 1.0f; memcpy(m.specular_color, om.specular, 12); m.hasReflective = 1.0f; }
            else { m.specular_exponent = 1.0f; memcpy(m.specular_color, om.specular, 12); m.hasReflective = 1.0f; m.hasRefractive = 1.0f; m.indexOfRefraction = om.ior; }
            memcpy(m.transmittance, om.transmittance, 12);
        } else { m.color[0] = m.color[1] = m.color[2] = 1.0f; }   // (the reference reads materials[i].transmittance out of bounds here)
        hs.offsets.push_back((int)hs.materials.size());
        hs.materials.push_back(m);
    }
    std::vector<float> tri18; std::vector<int> mtl;
    for (int i = 0; i < numShapes; i++) {
        const std::vector<int>& sh = obj.shapes[i];
        for (size_t j = 0; j + 2 < sh.size(); j += 3) {
            for (int k = 0; k < 3; k++) {
                int vi = sh[j + k];
                if (vi < 0 || (size_t)vi >= nv) { err = "OBJ face references vertex " + std::to_string(vi + 1) + " of " + std::to_string(nv); return false; }
            }
            size_t base = tri18.size();
            tri18.resize(base + 18);
            for (int k = 0; k < 3; k++)
                for (int c = 0; c < 3; c++) { tri18[base + 3 * k + c] = obj.v[3 * (size_t)sh[j + k] + c]; tri18[base + 9 + 3 * k + c] = normals[3 * (size_t)sh[j + k] + c]; }
            mtl.push_back(i);
        }
    }
    buildTree(tri18.data(), mtl.data(), (int)mtl.size(), kdDepth, hs);
    return true;
}

} // namespace

struct kdpt_host_scene { HostScene s; };

extern "C" {

int kdpt_scene_load(const char* scene_txt, const char* obj_path, const char* mtl_dir, int kd_depth, int synth_normals,
                    kdpt_host_scene** out)
{
    if (!scene_txt || !out) { kdpt::setError("kdpt_scene_load: null argument"); return KDPT_ERR_ARG; }
    *out = nullptr;
    kdpt_host_scene* h = new kdpt_host_scene();
    std::string err;
    if (!loadSceneTxt(scene_txt, h->s, err)) { kdpt::setError("kdpt_scene_load: " + err); delete h; return KDPT_ERR_ARG; }
    if (obj_path && obj_path[0]) {
        std::string dir = mtl_dir ? mtl_dir : "";
        if (!mtl_dir) { std::string p = obj_path; size_t k = p.find_last_of("/\\"); dir = k == std::string::npos ? "" : p.substr(0, k + 1); }
        if (!attachObj(obj_path, dir, kd_depth > 0 ? kd_depth : 13, synth_normals, h->s, err)) { kdpt::setError("kdpt_scene_load: " + err); delete h; return KDPT_ERR_ARG; }
    }
    *out = h;
    return KDPT_OK;
}

int kdpt_scene_from_triangles(const float* tri18, const int* mtl, int n, int kd_depth, kdpt_host_scene** out)
{
    if (!out || n < 0 || (n > 0 && !tri18)) { kdpt::setError("kdpt_scene_from_triangles: bad argument"); return KDPT_ERR_ARG; }
    kdpt_host_scene* h = new kdpt_host_scene();
    memset(&h->s.camera, 0, sizeof h->s.camera);
    kdpt_material m; memset(&m, 0, sizeof m); m.color[0] = m.color[1] = m.color[2] = 1.0f;
    h->s.materials.push_back(m); h->s.offsets.push_back(0); h->s.traceDepth = 8;
    buildTree(tri18, mtl, n, kd_depth > 0 ? kd_depth : 13, h->s);
    *out = h;
    return KDPT_OK;
}

int kdpt_scene_from_triangles_gpu(const float* tri18, const int* mtl, int n, int kd_depth, int device, float* device_ms,
                                  kdpt_host_scene** out)
{
    if (!out || n < 0 || (n > 0 && !tri18)) { kdpt::setError("kdpt_scene_from_triangles_gpu: bad argument"); return KDPT_ERR_ARG; }
    *out = nullptr;
    kdpt_host_scene* h = new kdpt_host_scene();
    memset(&h->s.camera, 0, sizeof h->s.camera);
    kdpt_material m; memset(&m, 0, sizeof m); m.color[0] = m.color[1] = m.color[2] = 1.0f;
    h->s.materials.push_back(m); h->s.offsets.push_back(0); h->s.traceDepth = 8;
    int rc = kdpt::buildTreeDevice(tri18, mtl, n, kd_depth > 0 ? kd_depth : 13, device, h->s.nodes, h->s.tris, h->s.triSource, device_ms);
    if (rc != KDPT_OK) { delete h; return rc; }
    *out = h;
    return KDPT_OK;
}

// ---- binary scene cache: the arrays above in one file, so that a 1 M / 10 M-triangle config starts
// without re-parsing the OBJ and re-building the tree. Layout: magic, version, 8 counts, camera, traceDepth, iterations,
// image name, the arrays back to back, then FNV-1a-64 of everything before it. Same-endianness, same-ABI file (the
// PODs are the reference's own layouts); a file that does not check out is refused, never half-loaded.
extern "C++" {
namespace {
const char kCacheMagic[8] = {'K', 'D', 'P', 'T', 'S', 'C', 'N', '1'};
uint64_t fnv1a(uint64_t h, const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
struct CacheWriter {
    FILE* f; uint64_t h = 1469598103934665603ull; bool ok = true;
    void put(const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) ok = false; h = fnv1a(h, p, n); }
    template <class V> void vec(const V& v) { put(v.data(), v.size() * sizeof(typename V::value_type)); }
};
struct CacheReader {
    FILE* f; uint64_t h = 1469598103934665603ull; bool ok = true;
    void get(void* p, size_t n) { if (n && fread(p, 1, n, f) != n) ok = false; else h = fnv1a(h, p, n); }
    template <class V> void vec(V& v, uint64_t count, uint64_t limit)
    {
        if (!ok || count > limit) { ok = false; return; }
        v.resize((size_t)count); get(v.data(), v.size() * sizeof(typename V::value_type));
    }
};
bool readWholeFile(const std::string& path, std::string& out)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16]; size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}
} // namespace
} // extern "C++"

int kdpt_scene_save(const kdpt_host_scene* hsc, const char* path)
{
    if (!hsc || !path) { kdpt::setError("kdpt_scene_save: null argument"); return KDPT_ERR_ARG; }
    const HostScene& s = hsc->s;
    const std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { kdpt::setError(std::string("kdpt_scene_save: cannot write ") + tmp); return KDPT_ERR_ARG; }
    CacheWriter w{f};
    w.put(kCacheMagic, 8);
    const uint64_t hdr[10] = {1, s.nodes.size(), s.tris.size(), s.triSource.size(), s.geoms.size(), s.materials.size(), s.offsets.size(),
                              (uint64_t)(int64_t)s.traceDepth, (uint64_t)(int64_t)s.iterations, s.imageName.size()};
    w.put(hdr, sizeof hdr);
    w.put(&s.camera, sizeof s.camera);
    w.put(s.imageName.data(), s.imageName.size());
    w.vec(s.nodes); w.vec(s.tris); w.vec(s.triSource); w.vec(s.geoms); w.vec(s.materials); w.vec(s.offsets);
    const uint64_t sum = w.h;
    if (fwrite(&sum, 1, 8, f) != 8) w.ok = false;
    if (fclose(f) != 0) w.ok = false;
    if (!w.ok || rename(tmp.c_str(), path) != 0) { remove(tmp.c_str()); kdpt::setError(std::string("kdpt_scene_save: write failed: ") + path); return KDPT_ERR_ARG; }
    return KDPT_OK;
}

int kdpt_scene_load_cache(const char* path, kdpt_host_scene** out)
{
    if (!path || !out) { kdpt::setError("kdpt_scene_load_cache: null argument"); return KDPT_ERR_ARG; }
    *out = nullptr;
    FILE* f = fopen(path, "rb");
    if (!f) { kdpt::setError(std::string("kdpt_scene_load_cache: cannot open ") + path); return KDPT_ERR_ARG; }
    fseek(f, 0, SEEK_END);
    const uint64_t fileBytes = (uint64_t)ftell(f);
    fseek(f, 0, SEEK_SET);
    kdpt_host_scene* h = new kdpt_host_scene();
    HostScene& s = h->s;
    CacheReader r{f};
    char magic[8]; uint64_t hdr[10] = {0};
    r.get(magic, 8); r.get(hdr, sizeof hdr);
    if (r.ok && (memcmp(magic, kCacheMagic, 8) != 0 || hdr[0] != 1)) r.ok = false;
    if (r.ok) {
        r.get(&s.camera, sizeof s.camera);
        s.traceDepth = (int)(int64_t)hdr[7]; s.iterations = (int)(int64_t)hdr[8];
        if (hdr[9] > 4096) r.ok = false;
        else { s.imageName.resize((size_t)hdr[9]); r.get(&s.imageName[0], s.imageName.size()); }
        // counts are bounded by the file size before anything is allocated
        r.vec(s.nodes, hdr[1], fileBytes / sizeof(kdpt_node_bare)); r.vec(s.tris, hdr[2], fileBytes / sizeof(kdpt_tri_bare));
        r.vec(s.triSource, hdr[3], fileBytes / sizeof(int)); r.vec(s.geoms, hdr[4], fileBytes / sizeof(kdpt_geom));
        r.vec(s.materials, hdr[5], fileBytes / sizeof(kdpt_material)); r.vec(s.offsets, hdr[6], fileBytes / sizeof(int));
    }
    uint64_t sum = 0;
    const uint64_t expect = r.h;
    if (r.ok && (fread(&sum, 1, 8, f) != 8 || sum != expect || fgetc(f) != EOF)) r.ok = false;
    fclose(f);
    if (!r.ok) { delete h; kdpt::setError(std::string("kdpt_scene_load_cache: not a valid scene cache (truncated, corrupt or another version): ") + path); return KDPT_ERR_ARG; }
    *out = h;
    return KDPT_OK;
}

int kdpt_scene_load_cached(const char* scene_txt, const char* obj_path, const char* mtl_dir, int kd_depth, int synth_normals,
                           const char* cache_dir, kdpt_host_scene** out, int* cache_hit)
{
    if (!scene_txt || !out || !cache_dir) { kdpt::setError("kdpt_scene_load_cached: null argument"); return KDPT_ERR_ARG; }
    if (cache_hit) *cache_hit = 0;
    // content address: the bytes of the scene file, the OBJ and every material library it names, plus the build options
    uint64_t key = 1469598103934665603ull;
    std::string bytes;
    if (!readWholeFile(scene_txt, bytes)) { kdpt::setError(std::string("kdpt_scene_load_cached: cannot read ") + scene_txt); return KDPT_ERR_ARG; }
    key = fnv1a(key, bytes.data(), bytes.size());
    std::string dir = mtl_dir ? mtl_dir : "";
    if (obj_path && obj_path[0]) {
        if (!mtl_dir) { std::string p = obj_path; size_t k = p.find_last_of("/\\"); dir = k == std::string::npos ? "" : p.substr(0, k + 1); }
        if (!readWholeFile(obj_path, bytes)) { kdpt::setError(std::string("kdpt_scene_load_cached: cannot read ") + obj_path); return KDPT_ERR_ARG; }
        key = fnv1a(key, "\0obj", 4); key = fnv1a(key, bytes.data(), bytes.size());
        std::istringstream in(bytes);
        std::string line, lib;
        while (safeGetline(in, line)) {
            if (line.compare(0, 7, "mtllib ") == 0 || line.compare(0, 7, "mtllib\t") == 0) {
                std::vector<std::string> t = tokenize(line);
                if (t.size() > 1 && readWholeFile(dir + t[1], lib)) { key = fnv1a(key, "\0mtl", 4); key = fnv1a(key, lib.data(), lib.size()); }
            }
        }
    }
    const int opts[3] = {kd_depth > 0 ? kd_depth : 13, synth_normals ? 1 : 0, (int)sizeof(kdpt_geom)};
    key = fnv1a(key, opts, sizeof opts);
    char name[64];
    snprintf(name, sizeof name, "kdpt-%016llx.scn", (unsigned long long)key);
    std::string path = std::string(cache_dir);
    if (!path.empty() && path.back() != '/') path += '/';
    path += name;
    if (kdpt_scene_load_cache(path.c_str(), out) == KDPT_OK) { if (cache_hit) *cache_hit = 1; return KDPT_OK; }
    int rc = kdpt_scene_load(scene_txt, obj_path, mtl_dir, kd_depth, synth_normals, out);
    if (rc != KDPT_OK) return rc;
    kdpt_scene_save(*out, path.c_str());          // best effort: an unwritable cache directory does not fail the load
    return KDPT_OK;
}

int kdpt_scene_view(const kdpt_host_scene* h, kdpt_scene* view, int* iterations)
{
    if (!h || !view) { kdpt::setError("kdpt_scene_view: null argument"); return KDPT_ERR_ARG; }
    const HostScene& s = h->s;
    view->nodes = s.nodes.data(); view->numNodes = (int)s.nodes.size();
    view->tris = s.tris.data(); view->numTris = (int)s.tris.size();
    view->geoms = s.geoms.data(); view->numGeoms = (int)s.geoms.size();
    view->materials = s.materials.data(); view->numMaterials = (int)s.materials.size();
    view->objMaterialOffsets = s.offsets.data(); view->numShapes = (int)s.offsets.size();
    view->camera = s.camera; view->traceDepth = s.traceDepth;
    if (iterations) *iterations = s.iterations;
    return KDPT_OK;
}

const int* kdpt_scene_triangle_sources(const kdpt_host_scene* h) { return h ? h->s.triSource.data() : nullptr; }

void kdpt_scene_free(kdpt_host_scene* h) { delete h; }

} // extern "C"
