// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_math.h header). Parity: unpinned by the reference (it ships no tests).
//
// CPU restatement of the Shi-Tomasi detector:
//   ShiTomasi::Extract / GetKeyPoints / IsLocalMaximum      modules/features/shi_tomasi.cc:38-54,75-160
//   ShiTomasi::FastSobelXYandScore / DetectCorner           modules/features/shi_tomasi.cc:163-409
// Two modes:
//   literal = 1  simulates the reference's single-pass row-pointer rotation as written, including what it leaves in
//                the first four / last four score rows (rows 2,3 are differentiated over image rows
//                {1,2,2},{2,2,3}; the first-row loops use `rows` as the COLUMN bound; score rows >= rows-3 are never
//                written and keep whatever the buffer held, including -1 marks of earlier calls);
//   literal = 0  the CLEAN definition the CUDA kernel implements: identical arithmetic wherever the reference's
//                gradients are aligned (score rows 4 .. rows-5, columns 1 .. cols-2), score 0 in the border band.
// tests/test_oracle_shi.py shows that both modes detect the same keypoints farther than 4 + 15 px from the border.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

namespace {

struct Shi {
  int nms = 7;
  unsigned next_id = 0;
  int rows = 0, cols = 0;
  std::vector<int16_t> gx, gy;
  std::vector<float> score;
};

inline float min_eig(float t0, float t1, float t2) {  // ComputeMinEigenValue, shi_tomasi.cc:396-405
  const float tr = t0 + t2;
  const float det = t0 * t2 - t1 * t1;
  const float root = tr * tr - 4 * det;
  return (float)((tr - sqrtf(root)) * 0.5);
}

// clean definition -------------------------------------------------------------------------------------------------
void scores_clean(Shi& S, const uint8_t* im, int pitch) {
  const int R = S.rows, C = S.cols;
  auto I = [&](int r, int c) { return (int)im[(size_t)r * pitch + c]; };
  // gradient "centered at image row r": gx(r,c), gy(r,c)
  auto GX = [&](int r, int c) -> int {
    if (c <= 0 || c >= C - 1) return 0;  // never written by the reference: stays 0
    return (I(r - 1, c + 1) + 2 * I(r, c + 1) + I(r + 1, c + 1)) - (I(r - 1, c - 1) + 2 * I(r, c - 1) + I(r + 1, c - 1));
  };
  auto Rw = [&](int r, int c) -> int {
    if (c == 0) return 2 * I(r, 0) + 2 * I(r, 1);
    if (c == C - 1) return 2 * I(r, C - 1) + 2 * I(r, C - 2);
    return I(r, c - 1) + 2 * I(r, c) + I(r, c + 1);
  };
  auto GY = [&](int r, int c) -> int { return (int)(int16_t)(Rw(r + 1, c) - Rw(r - 1, c)); };
  const float inv_size = 1.f / 9.f;
  for (int s = 0; s < R; s++)
    for (int c = 0; c < C; c++) {
      float v = 0.f;
      if (s >= 4 && s <= R - 5 && c >= 1 && c <= C - 2) {
        int g11 = 0, g12 = 0, g22 = 0;
        for (int dr = -1; dr <= 1; dr++)
          for (int dc = -1; dc <= 1; dc++) {
            const int x = (int16_t)GX(s + dr, c + dc), y = GY(s + dr, c + dc);
            g11 += x * x;
            g12 += x * y;
            g22 += y * y;
          }
        v = min_eig((float)g11 * inv_size, (float)g12 * inv_size, (float)g22 * inv_size);
      }
      S.score[(size_t)s * C + c] = v;
    }
}

// literal simulation ------------------------------------------------------------------------------------------------
void scores_literal(Shi& S, const uint8_t* im, int pitch) {
  const int rows_l = S.rows, cols_l = S.cols;
  auto IM = [&](int r) { return im + (size_t)r * pitch; };
  auto XG = [&](int r) { return S.gx.data() + (size_t)r * cols_l; };
  auto YG = [&](int r) { return S.gy.data() + (size_t)r * cols_l; };
  const uint8_t* pIm[3];
  int16_t *pX[3], *pY[3];
  std::vector<int16_t> r1(cols_l), r2(cols_l), r3(cols_l);
  int16_t c1, c2, c3;
  float G11a = 0, G12a = 0, G22a = 0, G11b = 0, G12b = 0, G22b = 0;
  const float inv_size = 1.f / 9.f;
  float* pScore = nullptr;
  auto detect = [&](int col) {  // DetectCorner, :347-394
    float t0, t1, t2;
    auto colsum = [&](int cc, float& a, float& b, float& c) {
      a = (float)(pX[0][cc] * pX[0][cc] + pX[1][cc] * pX[1][cc] + pX[2][cc] * pX[2][cc]);
      b = (float)(pX[0][cc] * pY[0][cc] + pX[1][cc] * pY[1][cc] + pX[2][cc] * pY[2][cc]);
      c = (float)(pY[0][cc] * pY[0][cc] + pY[1][cc] * pY[1][cc] + pY[2][cc] * pY[2][cc]);
    };
    if (col == 1) {
      colsum(col, G11a, G12a, G22a);
      colsum(col + 1, G11b, G12b, G22b);
      float a, b, c;
      colsum(col - 1, a, b, c);
      t0 = (G11a + G11b + a) * inv_size;
      t1 = (G12a + G12b + b) * inv_size;
      t2 = (G22a + G22b + c) * inv_size;
    } else {
      t0 = G11a + G11b;
      t1 = G12a + G12b;
      t2 = G22a + G22b;
      G11a = G11b;
      G12a = G12b;
      G22a = G22b;
      colsum(col + 1, G11b, G12b, G22b);
      t0 = (t0 + G11b) * inv_size;
      t1 = (t1 + G12b) * inv_size;
      t2 = (t2 + G22b) * inv_size;
    }
    pScore[col] = min_eig(t0, t1, t2);
  };
  // first row (:168-190) — note the column bound `rows_l`
  pX[1] = XG(0);
  pY[1] = YG(0);
  pIm[1] = IM(0);
  pIm[2] = IM(1);
  c1 = pIm[1][0] + pIm[1][0] + pIm[2][0] + pIm[1][0];
  c2 = pIm[1][1] + pIm[1][1] + pIm[2][1] + pIm[1][1];
  c3 = pIm[1][2] + pIm[1][2] + pIm[2][2] + pIm[1][2];
  pX[1][1] = c3 - c1;
  for (int j = 2; j < rows_l - 1 && j + 1 < cols_l; j++) {  // (guarded: the reference would run off a row when rows > cols)
    c1 = c2;
    c2 = c3;
    c3 = pIm[1][j + 1] + pIm[1][j + 1] + pIm[2][j + 1] + pIm[2][j + 1];
    pX[1][j] = c3 - c1;
  }
  // second row (:192-246)
  pX[2] = XG(1);
  pY[2] = YG(1);
  pIm[0] = pIm[1];
  pIm[1] = pIm[2];
  pIm[2] = IM(2);
  r1[0] = pIm[0][0] + pIm[0][0] + pIm[0][1] + pIm[0][1];
  r2[0] = pIm[1][0] + pIm[1][0] + pIm[1][1] + pIm[1][1];
  r3[0] = pIm[2][0] + pIm[2][0] + pIm[2][1] + pIm[2][1];
  pY[2][0] = r3[0] - r1[0];
  c1 = pIm[0][0] + pIm[1][0] + pIm[1][0] + pIm[2][0];
  c2 = pIm[0][1] + pIm[1][1] + pIm[1][1] + pIm[2][1];
  c3 = pIm[0][2] + pIm[1][2] + pIm[1][2] + pIm[2][2];
  pX[2][1] = c3 - c1;
  r1[1] = pIm[0][0] + pIm[0][1] + pIm[0][1] + pIm[2][2];  // sic (:223)
  r2[1] = pIm[1][0] + pIm[1][1] + pIm[1][1] + pIm[1][2];
  r3[1] = pIm[2][0] + pIm[2][1] + pIm[2][1] + pIm[2][2];
  pY[2][1] = r3[1] - r1[1];
  for (int j = 2; j < cols_l - 1; j++) {
    c1 = c2;
    c2 = c3;
    c3 = pIm[0][j + 1] + pIm[1][j + 1] + pIm[1][j + 1] + pIm[2][j + 1];
    pX[2][j] = c3 - c1;
    r1[j] = pIm[0][j - 1] + pIm[0][j] + pIm[0][j] + pIm[0][j + 1];
    r2[j] = pIm[1][j - 1] + pIm[1][j] + pIm[1][j] + pIm[1][j + 1];
    r3[j] = pIm[2][j - 1] + pIm[2][j] + pIm[2][j] + pIm[2][j + 1];
    pY[2][j] = r3[j] - r1[j];
  }
  r1[cols_l - 1] = pIm[0][cols_l - 1] + pIm[0][cols_l - 1] + pIm[0][cols_l - 2] + pIm[0][cols_l - 2];
  r2[cols_l - 1] = pIm[1][cols_l - 1] + pIm[1][cols_l - 1] + pIm[1][cols_l - 2] + pIm[1][cols_l - 2];
  r3[cols_l - 1] = pIm[2][cols_l - 1] + pIm[2][cols_l - 1] + pIm[2][cols_l - 2] + pIm[2][cols_l - 2];
  pY[2][cols_l - 1] = r3[cols_l - 1] - r1[cols_l - 1];
  // inner rows (:248-309)
  int i = 2;
  for (; i < rows_l - 1; i++) {
    pX[0] = pX[1];
    pX[1] = pX[2];
    pX[2] = XG(i);
    pY[0] = pY[1];
    pY[1] = pY[2];
    pY[2] = YG(i);
    pIm[0] = pIm[1];
    pIm[1] = pIm[2];
    pIm[2] = IM(i);
    r1 = r2;
    r2 = r3;
    r3[0] = pIm[2][0] + pIm[2][0] + pIm[2][1] + pIm[2][1];
    pY[2][0] = r3[0] - r1[0];
    c1 = pIm[0][0] + pIm[1][0] + pIm[1][0] + pIm[2][0];
    c2 = pIm[0][1] + pIm[1][1] + pIm[1][1] + pIm[2][1];
    c3 = pIm[0][2] + pIm[1][2] + pIm[1][2] + pIm[2][2];
    pX[2][1] = c3 - c1;
    r3[1] = pIm[2][0] + pIm[2][1] + pIm[2][1] + pIm[2][2];
    pY[2][1] = r3[1] - r1[1];
    pScore = S.score.data() + (size_t)(i - 2) * cols_l;
    for (int j = 2; j < cols_l - 1; j++) {
      c1 = c2;
      c2 = c3;
      c3 = pIm[0][j + 1] + pIm[1][j + 1] + pIm[1][j + 1] + pIm[2][j + 1];
      pX[2][j] = c3 - c1;
      r3[j] = pIm[2][j - 1] + pIm[2][j] + pIm[2][j] + pIm[2][j + 1];
      pY[2][j] = r3[j] - r1[j];
      detect(j - 1);
    }
    r3[cols_l - 1] = pIm[2][cols_l - 1] + pIm[2][cols_l - 1] + pIm[2][cols_l - 2] + pIm[2][cols_l - 2];
    pY[2][cols_l - 1] = r3[cols_l - 1] - r1[cols_l - 1];
    detect(cols_l - 2);
  }
  // last row (:311-338): stale image pointers, score row pointer still the last inner one
  pX[0] = XG(rows_l - 3);
  pX[1] = XG(rows_l - 2);
  pX[2] = XG(rows_l - 1);
  pY[0] = YG(rows_l - 3);
  pY[1] = YG(rows_l - 2);
  pY[2] = YG(rows_l - 1);
  c1 = pIm[1][0] + pIm[1][0] + pIm[2][0] + pIm[1][0];
  c2 = pIm[1][1] + pIm[1][1] + pIm[2][1] + pIm[1][1];
  c3 = pIm[1][2] + pIm[1][2] + pIm[2][2] + pIm[1][2];
  pX[2][1] = c3 - c1;
  for (int j = 1; j < rows_l - 1 && j + 1 < cols_l; j++) {
    c1 = c2;
    c2 = c3;
    c3 = pIm[0][j + 1] + pIm[1][j + 1] + pIm[1][j + 1] + pIm[2][j + 1];
    pX[2][j] = c3 - c1;
    detect(j);
  }
}

bool is_local_max(const Shi& S, int r, int c) {  // :123-160
  const int NnoPrev = S.nms, NPrev = 15;
  const int nrows = S.rows, ncols = S.cols;
  const int minRow = r - NPrev < 0 ? 0 : r - NPrev, minCol = c - NPrev < 0 ? 0 : c - NPrev;
  const int maxRow = r + NPrev > nrows - 1 ? nrows - 1 : r + NPrev, maxCol = c + NPrev > ncols - 1 ? ncols - 1 : c + NPrev;
  const int minRi = r - NnoPrev < 0 ? 0 : r - NnoPrev, minCi = c - NnoPrev < 0 ? 0 : c - NnoPrev;
  const int maxRi = r + NnoPrev > nrows - 1 ? nrows - 1 : r + NnoPrev, maxCi = c + NnoPrev > ncols - 1 ? ncols - 1 : c + NnoPrev;
  const float cur = S.score[(size_t)r * ncols + c];
  if (cur == -1.f) return false;
  if (cur < 80) return false;
  for (int i = minRow; i <= maxRow; i++)
    for (int j = minCol; j <= maxCol; j++) {
      const float v = S.score[(size_t)i * ncols + j];
      if (v == -1.f) return false;
      if (i >= minRi && i <= maxRi && j >= minCi && j <= maxCi && v > cur) return false;
    }
  return true;
}

}  // namespace

extern "C" {

void* orc_shi_create(int nms_window) {
  Shi* s = new Shi();
  s->nms = nms_window;
  return s;
}
void orc_shi_destroy(void* p) { delete static_cast<Shi*>(p); }

// Extract (:38-54). existing: keypoints already in the frame (their rounded pixel gets score -1). Output: the NEW
// keypoints in raster order with class ids from the running counter. Returns the count (or -1 on bad arguments).
int orc_shi_extract(void* p, const uint8_t* im, int w, int h, int pitch, const float* existing, int n_existing,
                    int literal, float* out_xy, int32_t* out_id, int capacity, float* scores_out) {
  Shi& S = *static_cast<Shi*>(p);
  if (w < 8 || h < 8) return -1;
  if (S.rows != h || S.cols != w) {  // ResizeBuffers (:56-67)
    S.rows = h;
    S.cols = w;
    S.gx.assign((size_t)w * h, 0);
    S.gy.assign((size_t)w * h, 0);
    S.score.assign((size_t)w * h, 0.f);
  }
  if (literal)
    scores_literal(S, im, pitch);
  else
    scores_clean(S, im, pitch);
  for (int i = 0; i < n_existing; i++) {  // :92-96 (out-of-range coordinates are undefined behaviour there; skipped here)
    const int r = (int)round((double)existing[2 * i + 1]), c = (int)round((double)existing[2 * i]);
    if (r >= 0 && r < h && c >= 0 && c < w) S.score[(size_t)r * w + c] = -1.f;
  }
  if (scores_out) memcpy(scores_out, S.score.data(), (size_t)w * h * sizeof(float));
  int n = 0;
  for (int r = 0; r < h; r++)
    for (int c = 0; c < w; c++)
      if (is_local_max(S, r, c)) {
        if (n < capacity) {
          out_xy[2 * n] = (float)c;
          out_xy[2 * n + 1] = (float)r;
          out_id[n] = (int32_t)S.next_id;
        }
        S.next_id++;
        n++;
      }
  return n;
}

}  // extern "C"
