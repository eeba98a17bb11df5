// ORACLE — TEST INFRASTRUCTURE ONLY (pinned against oracle/_ref, see vina_oracle.hpp).
// CPU restatement of the VINA-SLAM per-scan hot path; each function cites the
// reference file:line it follows.
#include "vina_oracle.hpp"
#include <chrono>
#include <cstdio>
#include <thread>

namespace vo
{
static inline double now_s()
{
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------------------
// src/core/point_utils.cpp:3-34
void calcBodyVar(Vec3& pb, const float range_inc, const float degree_inc, Mat3& var)
{
  if (pb[2] == 0) pb[2] = 0.0001;

  float range = std::sqrt(pb[0] * pb[0] + pb[1] * pb[1] + pb[2] * pb[2]);
  float range_var = range_inc * range_inc;

  double sd = std::sin((degree_inc)*M_PI / 180.0);
  double dv = std::pow(sd, 2);
  Mat<2, 2> direction_var = Mat<2, 2>::Zero();
  direction_var(0, 0) = dv;
  direction_var(1, 1) = dv;

  Vec3 direction = pb;
  normalize(direction);
  Mat3 direction_hat = hat(direction);

  Vec3 base_vector1 = V3(1, 1, -(direction[0] + direction[1]) / direction[2]);
  normalize(base_vector1);
  Vec3 base_vector2 = cross(base_vector1, direction);
  normalize(base_vector2);

  Mat<3, 2> N;
  for (int i = 0; i < 3; i++)
  {
    N(i, 0) = base_vector1[i];
    N(i, 1) = base_vector2[i];
  }
  Mat<3, 2> A = ((double)range * direction_hat) * N;
  var = (direction * (double)range_var) * direction.transpose() + (A * direction_var) * A.transpose();
}

// src/core/point_utils.cpp:36-52
void var_init(IMUST& ext, Cloud& pl_cur, PVecPtr pptr, double dept_err, double beam_err)
{
  int plsize = pl_cur.size();
  pptr->clear();
  pptr->resize(plsize);
  for (int i = 0; i < plsize; i++)
  {
    PointXYZT& ap = pl_cur[i];
    pointVar& pv = pptr->at(i);
    pv.pnt = V3(ap.x, ap.y, ap.z);
    calcBodyVar(pv.pnt, dept_err, beam_err, pv.var);
    pv.pnt = ext.R * pv.pnt + ext.p;
    pv.var = ext.R * pv.var * ext.R.transpose();
  }
}

// src/core/point_utils.cpp:54-65
void pvec_update(PVecPtr pptr, IMUST& x_curr, std::vector<Vec3>& pwld)
{
  Mat3 rot_var = x_curr.cov.block<3, 3>(0, 0);
  Mat3 tsl_var = x_curr.cov.block<3, 3>(3, 3);
  for (pointVar& pv : *pptr)
  {
    Mat3 phat = hat(pv.pnt);
    pv.var = x_curr.R * pv.var * x_curr.R.transpose() + phat * rot_var * phat.transpose() + tsl_var;
    pwld.push_back(x_curr.R * pv.pnt + x_curr.p);
  }
}

// include/vina_slam/core/point_utils.hpp:7-44
void down_sampling_voxel(Cloud& pl_feat, double voxel_size)
{
  if (voxel_size < 0.001) return;
  std::unordered_map<VOXEL_LOC, PointXYZT, VoxelHash> feat_map;
  float loc_xyz[3];
  for (PointXYZT& p_c : pl_feat)
  {
    const float data[3] = { p_c.x, p_c.y, p_c.z };
    for (int j = 0; j < 3; j++)
    {
      loc_xyz[j] = data[j] / voxel_size;
      if (loc_xyz[j] < 0) loc_xyz[j] -= 1.0;
    }
    VOXEL_LOC position((int64_t)loc_xyz[0], (int64_t)loc_xyz[1], (int64_t)loc_xyz[2]);
    auto iter = feat_map.find(position);
    if (iter == feat_map.end())
    {
      PointXYZT pp = p_c;
      pp.curvature = 1;
      feat_map[position] = pp;
    }
    else
    {
      PointXYZT& pp = iter->second;
      pp.x = (pp.x * pp.curvature + p_c.x) / (pp.curvature + 1);
      pp.y = (pp.y * pp.curvature + p_c.y) / (pp.curvature + 1);
      pp.z = (pp.z * pp.curvature + p_c.z) / (pp.curvature + 1);
      pp.curvature += 1;
    }
  }
  pl_feat.clear();
  for (auto iter = feat_map.begin(); iter != feat_map.end(); ++iter) pl_feat.push_back(iter->second);
}

// src/mapping/voxel_map.cpp:56-65 (== 246-253): double divide -> float ->
// "-1 if negative" in float -> truncate
VOXEL_LOC voxel_key(const Vec3& pw, double voxel_size)
{
  float loc[3];
  for (int j = 0; j < 3; j++)
  {
    loc[j] = pw[j] / voxel_size;
    if (loc[j] < 0) loc[j] -= 1;
  }
  return VOXEL_LOC(loc[0], loc[1], loc[2]);
}

// ---------------------------------------------------------------------------
// src/mapping/octree.cpp:83-92
void Bf_var(const pointVar& pv, Mat9& bcov, const Vec3& vec)
{
  Mat<6, 3> Bi = Mat<6, 3>::Zero();
  Bi(0, 0) = 2 * vec[0];
  Bi(1, 0) = vec[1];
  Bi(1, 1) = vec[0];
  Bi(2, 0) = vec[2];
  Bi(2, 2) = vec[0];
  Bi(3, 1) = 2 * vec[1];
  Bi(4, 1) = vec[2];
  Bi(4, 2) = vec[1];
  Bi(5, 2) = 2 * vec[2];
  Mat<6, 3> Biup = Bi * pv.var;
  bcov.setBlock<6, 6>(0, 0, Biup * Bi.transpose());
  bcov.setBlock<6, 3>(0, 6, Biup);
  bcov.setBlock<3, 6>(6, 0, Biup.transpose());
  bcov.setBlock<3, 3>(6, 6, pv.var);
}

// src/mapping/octree.cpp:144-149
OctoTree::OctoTree(Globals* g, int _l, int _w) : G(g), layer(_l), octo_state(0), wdsize(_w)
{
  for (int i = 0; i < 8; i++) leaves[i] = nullptr;
  cov_add.setZero();
  eig_value.setZero();
  eig_vector.setZero();
  voxel_center[0] = voxel_center[1] = voxel_center[2] = 0;
  quater_length = 0;
}

// src/mapping/octree.cpp:151-177
void OctoTree::push(int ord, const pointVar& pv, const Vec3& pw, std::vector<SlideWindow*>& sws)
{
  if (sw == nullptr)
  {
    if (sws.size() != 0)
    {
      sw = sws.back();
      sws.pop_back();
      sw->resize(wdsize);
    }
    else
      sw = new SlideWindow(wdsize);
  }
  if (!isexist) isexist = true;

  int mord = G->mp[ord];
  if (layer < G->max_layer) sw->points[mord].push_back(pv);
  sw->pcrs_local[mord].push(pv.pnt);
  pcr_add.push(pw);
  Mat9 Bi;
  Bf_var(pv, Bi, pw);
  cov_add += Bi;
}

// src/mapping/octree.cpp:179-188
void OctoTree::push_fix(pointVar& pv)
{
  if (layer < G->max_layer) point_fix.push_back(pv);
  pcr_fix.push(pv.pnt);
  pcr_add.push(pv.pnt);
  Mat9 Bi;
  Bf_var(pv, Bi, pv.pnt);
  cov_add += Bi;
}

// src/mapping/octree.cpp:198-201
bool OctoTree::plane_judge(Vec3& eig_values)
{
  return (eig_values[0] < G->min_eigen_value && (eig_values[0] / eig_values[2]) < G->plane_eigen_value_thre[layer]);
}

// child creation shared by allocate/fix_divide/subdivide (octree.cpp:217-224, 266-273, 289-296)
OctoTree* OctoTree::make_child(int leafnum, const int xyz[3])
{
  OctoTree* c = new OctoTree(G, layer + 1, wdsize);
  c->voxel_center[0] = voxel_center[0] + (2 * xyz[0] - 1) * quater_length;
  c->voxel_center[1] = voxel_center[1] + (2 * xyz[1] - 1) * quater_length;
  c->voxel_center[2] = voxel_center[2] + (2 * xyz[2] - 1) * quater_length;
  c->quater_length = quater_length / 2;
  c->root_key = root_key;
  c->path = path | (leafnum << (3 * layer));
  leaves[leafnum] = c;
  return c;
}

// src/mapping/octree.cpp:203-228
void OctoTree::allocate(int ord, const pointVar& pv, const Vec3& pw, std::vector<SlideWindow*>& sws)
{
  if (octo_state == 0)
  {
    push(ord, pv, pw, sws);
  }
  else
  {
    int xyz[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++)
      if (pw[k] > voxel_center[k]) xyz[k] = 1;
    int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
    if (leaves[leafnum] == nullptr) make_child(leafnum, xyz);
    leaves[leafnum]->allocate(ord, pv, pw, sws);
  }
}

// src/mapping/octree.cpp:257-277
void OctoTree::fix_divide(std::vector<SlideWindow*>& sws)
{
  for (pointVar& pv : point_fix)
  {
    int xyz[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++)
      if (pv.pnt[k] > voxel_center[k]) xyz[k] = 1;
    int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
    if (leaves[leafnum] == nullptr) make_child(leafnum, xyz);
    leaves[leafnum]->push_fix(pv);
  }
}

// src/mapping/octree.cpp:279-300
void OctoTree::subdivide(int si, IMUST& xx, std::vector<SlideWindow*>& sws)
{
  for (pointVar& pv : sw->points[G->mp[si]])
  {
    Vec3 pw = xx.R * pv.pnt + xx.p;
    int xyz[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++)
      if (pw[k] > voxel_center[k]) xyz[k] = 1;
    int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
    if (leaves[leafnum] == nullptr) make_child(leafnum, xyz);
    leaves[leafnum]->push(si, pv, pw, sws);
  }
}

// src/mapping/octree.cpp:302-333
void OctoTree::plane_update()
{
  plane.center = pcr_add.v / (double)pcr_add.N;
  int l = 0;
  Vec3 u[3] = { eig_vector.col(0), eig_vector.col(1), eig_vector.col(2) };
  double nv = 1.0 / pcr_add.N;

  Mat<3, 9> u_c = Mat<3, 9>::Zero();
  for (int k = 0; k < 3; k++)
    if (k != l)
    {
      Mat3 ukl = u[k] * u[l].transpose();
      Mat<1, 9> fkl;
      fkl[0] = ukl(0, 0);
      fkl[1] = ukl(1, 0) + ukl(0, 1);
      fkl[2] = ukl(2, 0) + ukl(0, 2);
      fkl[3] = ukl(1, 1);
      fkl[4] = ukl(1, 2) + ukl(2, 1);
      fkl[5] = ukl(2, 2);
      Vec3 tail = -(dot(u[k], plane.center) * u[l] + dot(u[l], plane.center) * u[k]);
      fkl[6] = tail[0];
      fkl[7] = tail[1];
      fkl[8] = tail[2];
      u_c += ((nv / (eig_value[l] - eig_value[k])) * u[k]) * fkl;
    }

  Mat<3, 9> Jc = u_c * cov_add;
  plane.plane_var.setBlock<3, 3>(0, 0, Jc * u_c.transpose());
  Mat3 Jc_N = nv * Jc.block<3, 3>(0, 6);
  plane.plane_var.setBlock<3, 3>(0, 3, Jc_N);
  plane.plane_var.setBlock<3, 3>(3, 0, Jc_N.transpose());
  plane.plane_var.setBlock<3, 3>(3, 3, (nv * nv) * cov_add.block<3, 3>(6, 6));
  plane.normal = u[0];
  plane.radius = eig_value[2];
}

// src/mapping/octree.cpp:335-393
void OctoTree::recut(int win_count, std::vector<IMUST>& x_buf, std::vector<SlideWindow*>& sws)
{
  if (octo_state == 0)
  {
    if (layer >= 0)
    {
      opt_state = -1;
      if (pcr_add.N <= G->min_point[layer])
      {
        plane.is_plane = false;
        return;
      }
      if (!isexist || sw == nullptr) return;

      SelfAdjointEigen3 saes(pcr_add.cov());
      eig_value = saes.values;
      eig_vector = saes.vectors;
      plane.is_plane = plane_judge(eig_value);

      if (plane.is_plane)
        return;
      else if (layer >= G->max_layer)
        return;
    }

    if (pcr_fix.N != 0)
    {
      fix_divide(sws);
      PVec().swap(point_fix);
    }

    for (int i = 0; i < win_count; i++) subdivide(i, x_buf[i], sws);

    sw->clear();
    sws.push_back(sw);
    sw = nullptr;
    octo_state = 1;
  }

  for (int i = 0; i < 8; i++)
    if (leaves[i] != nullptr) leaves[i]->recut(win_count, x_buf, sws);
}

// src/mapping/octree.cpp:395-495. vox_opt is the LidarFactor filled by
// tras_opt (octree.cpp:498-521). With BA disabled (if_BA: 0; BA is out of
// scope) the entries tras_opt pushed are verbatim copies
// of this node's own pcr_add / eig_value / eig_vector, so "opt_state >= 0"
// keeps them and transforms only the marginalised frame (octree.cpp:410-422).
void OctoTree::margi(int win_count, int mgsize, std::vector<IMUST>& x_buf)
{
  if (octo_state == 0 && layer >= 0)
  {
    if (!isexist || sw == nullptr) return;
    std::vector<PointCluster> pcrs_world(wdsize);

    if (opt_state >= 0)
    {
      opt_state = -1;
      for (int i = 0; i < mgsize; i++)
        if (sw->pcrs_local[G->mp[i]].N != 0) pcrs_world[i].transform(sw->pcrs_local[G->mp[i]], x_buf[i]);
    }
    else
    {
      pcr_add = pcr_fix;
      for (int i = 0; i < win_count; i++)
        if (sw->pcrs_local[G->mp[i]].N != 0)
        {
          pcrs_world[i].transform(sw->pcrs_local[G->mp[i]], x_buf[i]);
          pcr_add += pcrs_world[i];
        }
      if (plane.is_plane)
      {
        SelfAdjointEigen3 saes(pcr_add.cov());
        eig_value = saes.values;
        eig_vector = saes.vectors;
      }
    }

    if (pcr_fix.N < G->max_points && plane.is_plane)
      if (pcr_add.N - last_num >= 5 || last_num <= 10)
      {
        plane_update();
        last_num = pcr_add.N;
      }

    if (pcr_fix.N < G->max_points)
    {
      for (int i = 0; i < mgsize; i++)
        if (pcrs_world[i].N != 0)
        {
          pcr_fix += pcrs_world[i];
          for (pointVar pv : sw->points[G->mp[i]])
          {
            pv.pnt = x_buf[i].R * pv.pnt + x_buf[i].p;
            point_fix.push_back(pv);
          }
        }
    }
    else
    {
      for (int i = 0; i < mgsize; i++)
        if (pcrs_world[i].N != 0) pcr_add -= pcrs_world[i];
      if (point_fix.size() != 0) PVec().swap(point_fix);
    }

    for (int i = 0; i < mgsize; i++)
      if (sw->pcrs_local[G->mp[i]].N != 0)
      {
        sw->pcrs_local[G->mp[i]].clear();
        sw->points[G->mp[i]].clear();
      }

    if (pcr_fix.N >= pcr_add.N)
      isexist = false;
    else
      isexist = true;
  }
  else
  {
    isexist = false;
    for (int i = 0; i < 8; i++)
      if (leaves[i] != nullptr)
      {
        leaves[i]->margi(win_count, mgsize, x_buf);
        isexist = isexist || leaves[i]->isexist;
      }
  }
}

// src/mapping/octree.cpp:498-521: which leaves become BA factors (opt_state >= 0)
static void tras_opt_mark(OctoTree* n, int& counter)
{
  if (n->octo_state == 0)
  {
    if (n->layer >= 0 && n->isexist && n->plane.is_plane && n->sw != nullptr)
    {
      if (n->eig_value[0] / n->eig_value[1] > 0.12) return;
      n->opt_state = counter++;
    }
  }
  else
  {
    for (int i = 0; i < 8; i++)
      if (n->leaves[i] != nullptr) tras_opt_mark(n->leaves[i], counter);
  }
}

// src/mapping/octree.cpp:498-521 with the container (the BA probe): same traversal, same acceptance test
void tras_opt_collect(OctoTree* n, LidarFactor& vox_opt, std::vector<OctoTree*>* nodes = nullptr)
{
  if (n->octo_state == 0)
  {
    if (n->layer >= 0 && n->isexist && n->plane.is_plane && n->sw != nullptr)
    {
      if (n->eig_value[0] / n->eig_value[1] > 0.12) return;
      double coe = 1;
      std::vector<PointCluster> pcrs(n->wdsize);
      for (int i = 0; i < n->wdsize; i++) pcrs[i] = n->sw->pcrs_local[n->G->mp[i]];
      vox_opt.push_voxel(pcrs, n->pcr_fix, coe, n->eig_value, n->eig_vector, n->pcr_add);
      if (nodes) nodes->push_back(n);
    }
  }
  else
  {
    for (int i = 0; i < 8; i++)
      if (n->leaves[i] != nullptr) tras_opt_collect(n->leaves[i], vox_opt, nodes);
  }
}

// ---------------------------------------------------------------------------
// src/mapping/factors.cpp:7-168 - the LiDAR BA factor. Expressions keep the reference's grouping and the eager
// left-to-right product order of omat.hpp, which is also what the reference build (oracle/_ref: factors.cpp
// compiled against ref_shim) evaluates - the two agree bit for bit in the strict builds.
void LidarFactor::push_voxel(std::vector<PointCluster>& vec_orig, PointCluster& fix, double coe, Vec3& eig_value,
                             Mat3& eig_vector, PointCluster& pcr_add)
{
  plvec_voxels.push_back(vec_orig);
  sig_vecs.push_back(fix);
  coeffs.push_back(coe);
  eig_values.push_back(eig_value);
  eig_vectors.push_back(eig_vector);
  pcr_adds.push_back(pcr_add);
}
void LidarFactor::clear()
{
  sig_vecs.clear();
  plvec_voxels.clear();
  eig_values.clear();
  eig_vectors.clear();
  pcr_adds.clear();
  coeffs.clear();
}

namespace
{
inline Mat3 outer(const Vec3& a, const Vec3& b)  // a * b.transpose()
{
  Mat3 r;
  for (int j = 0; j < 3; j++)
    for (int i = 0; i < 3; i++) r(i, j) = a[i] * b[j];
  return r;
}
}  // namespace

void LidarFactor::acc_evaluate2(const std::vector<IMUST>& xs, int head, int end, std::vector<double>& Hess,
                                std::vector<double>& JacT, double& residual)
{
  const int dim = 6 * win_size;
  Hess.assign((size_t)dim * dim, 0.0);
  JacT.assign(dim, 0.0);
  residual = 0;
  const int kk = 0;
  auto H = [&](int r, int c) -> double& { return Hess[r + (size_t)dim * c]; };
  std::vector<Vec3> viRiTuk(win_size);
  std::vector<Mat3> viRiTukukT(win_size);
  std::vector<Mat<3, 6>> Auk(win_size);
  const Mat3 I33 = Mat3::Identity();

  for (int a = head; a < end; a++)
  {
    std::vector<PointCluster>& sig_orig = plvec_voxels[a];
    double coe = coeffs[a];
    Vec3 lmbd = eig_values[a];
    Mat3 U = eig_vectors[a];
    int NN = pcr_adds[a].N;
    Vec3 vBar = pcr_adds[a].v / (double)NN;
    Vec3 u[3] = { U.col(0), U.col(1), U.col(2) };
    Vec3& uk = u[kk];
    Mat3 ukukT = outer(uk, uk);
    Mat3 umumT = Mat3::Zero();
    for (int i = 0; i < 3; i++)
      if (i != kk) umumT += (2.0 / (lmbd[kk] - lmbd[i])) * u[i] * u[i].transpose();

    for (int i = 0; i < win_size; i++)
      if (sig_orig[i].N != 0)
      {
        Mat3 Pi = sig_orig[i].P;
        Vec3 vi = sig_orig[i].v;
        Mat3 Ri = xs[i].R;
        double ni = sig_orig[i].N;
        Mat3 vihat = hat(vi);
        Vec3 RiTuk = Ri.transpose() * uk;
        Mat3 RiTukhat = hat(RiTuk);
        Vec3 PiRiTuk = Pi * RiTuk;
        viRiTuk[i] = vihat * RiTuk;
        viRiTukukT[i] = viRiTuk[i] * uk.transpose();
        Vec3 ti_v = xs[i].p - vBar;
        double ukTti_v = dot(uk, ti_v);
        Mat3 combo1 = hat(PiRiTuk) + vihat * ukTti_v;
        Vec3 combo2 = Ri * vi + ni * ti_v;
        Mat3 blk0 = (Ri * Pi + ti_v * vi.transpose()) * RiTukhat - Ri * combo1;
        Mat3 blk1 = combo2 * uk.transpose() + dot(combo2, uk) * I33;
        Auk[i].setBlock<3, 3>(0, 0, blk0);
        Auk[i].setBlock<3, 3>(0, 3, blk1);
        Auk[i] = Auk[i] / (double)NN;

        Vec6 jjt = Auk[i].transpose() * uk;
        for (int q = 0; q < 6; q++) JacT[6 * i + q] = JacT[6 * i + q] + (coe * jjt)[q];

        Mat3 HRt = (2.0 / NN * (1.0 - ni / NN)) * viRiTukukT[i];
        Mat6 Hb = Auk[i].transpose() * umumT * Auk[i];
        Vec3 jj3 = jjt.block<3, 1>(0, 0);
        Mat3 add00 = (2.0 / NN) * (combo1 - RiTukhat * Pi) * RiTukhat - (2.0 / NN / NN) * viRiTuk[i] * viRiTuk[i].transpose() -
                     0.5 * hat(jj3);
        Hb.setBlock<3, 3>(0, 0, Hb.block<3, 3>(0, 0) + add00);
        Hb.setBlock<3, 3>(0, 3, Hb.block<3, 3>(0, 3) + HRt);
        Hb.setBlock<3, 3>(3, 0, Hb.block<3, 3>(3, 0) + HRt.transpose());
        Hb.setBlock<3, 3>(3, 3, Hb.block<3, 3>(3, 3) + (2.0 / NN * (ni - ni * ni / NN)) * ukukT);
        Mat6 cH = coe * Hb;
        for (int c = 0; c < 6; c++)
          for (int r = 0; r < 6; r++) H(6 * i + r, 6 * i + c) = H(6 * i + r, 6 * i + c) + cH(r, c);
      }

    for (int i = 0; i < win_size - 1; i++)
      if (sig_orig[i].N != 0)
      {
        double ni = sig_orig[i].N;
        for (int j = i + 1; j < win_size; j++)
          if (sig_orig[j].N != 0)
          {
            double nj = sig_orig[j].N;
            Mat6 Hb = Auk[i].transpose() * umumT * Auk[j];
            Hb.setBlock<3, 3>(0, 0, Hb.block<3, 3>(0, 0) + (-2.0 / NN / NN) * viRiTuk[i] * viRiTuk[j].transpose());
            Hb.setBlock<3, 3>(0, 3, Hb.block<3, 3>(0, 3) + (-2.0 * nj / NN / NN) * viRiTukukT[i]);
            Hb.setBlock<3, 3>(3, 0, Hb.block<3, 3>(3, 0) + (-2.0 * ni / NN / NN) * viRiTukukT[j].transpose());
            Hb.setBlock<3, 3>(3, 3, Hb.block<3, 3>(3, 3) + (-2.0 * ni * nj / NN / NN) * ukukT);
            Mat6 cH = coe * Hb;
            for (int c = 0; c < 6; c++)
              for (int r = 0; r < 6; r++) H(6 * i + r, 6 * j + c) = H(6 * i + r, 6 * j + c) + cH(r, c);
          }
      }
    residual += coe * lmbd[kk];
  }
  for (int i = 1; i < win_size; i++)
    for (int j = 0; j < i; j++)
      for (int c = 0; c < 6; c++)
        for (int r = 0; r < 6; r++) H(6 * i + r, 6 * j + c) = H(6 * j + c, 6 * i + r);
}

void LidarFactor::evaluate_only_residual(const std::vector<IMUST>& xs, int head, int end, double& residual)
{
  residual = 0;
  int kk = 0;
  PointCluster pcr;
  for (int a = head; a < end; a++)
  {
    const std::vector<PointCluster>& sig_orig = plvec_voxels[a];
    PointCluster sig = sig_vecs[a];
    for (int i = 0; i < win_size; i++)
      if (sig_orig[i].N != 0)
      {
        pcr.transform(sig_orig[i], xs[i]);
        sig += pcr;
      }
    Vec3 vBar = sig.v / (double)sig.N;
    Mat3 cov;
    for (int j = 0; j < 3; j++)
      for (int i = 0; i < 3; i++) cov(i, j) = sig.P(i, j) / (double)sig.N - vBar[i] * vBar[j];
    SelfAdjointEigen3 saes(cov);
    eig_values[a] = saes.values;
    eig_vectors[a] = saes.vectors;
    pcr_adds[a] = sig;
    residual += coeffs[a] * saes.values[kk];
  }
}

// src/mapping/octree.cpp:551-595
int OctoTree::match(Vec3& wld, Plane*& pla, double& max_prob, Mat3& var_wld, double& sigma_d, OctoTree*& oc)
{
  int flag = 0;
  if (octo_state == 0)
  {
    if (plane.is_plane)
    {
      float dis_to_plane = std::fabs(dot(plane.normal, wld - plane.center));
      float dis_to_center = squaredNorm(plane.center - wld);
      float range_dis = (dis_to_center - dis_to_plane * dis_to_plane);
      if (range_dis <= 3 * 3 * plane.radius)
      {
        Mat<1, 6> J_nq;
        Vec3 d = wld - plane.center;
        for (int k = 0; k < 3; k++)
        {
          J_nq[k] = d[k];
          J_nq[3 + k] = -plane.normal[k];
        }
        Mat<1, 1> s1 = J_nq * plane.plane_var * J_nq.transpose();
        double sigma_l = s1[0];
        Mat<1, 1> s2 = plane.normal.transpose() * var_wld * plane.normal;
        sigma_l += s2[0];
        if (dis_to_plane < 3 * std::sqrt(sigma_l))
        {
          oc = this;
          sigma_d = sigma_l;
          pla = &plane;
          flag = 1;
        }
      }
    }
  }
  else
  {
    int xyz[3] = { 0, 0, 0 };
    for (int k = 0; k < 3; k++)
      if (wld[k] > voxel_center[k]) xyz[k] = 1;
    int leafnum = 4 * xyz[0] + 2 * xyz[1] + xyz[2];
    if (leaves[leafnum] != nullptr) flag = leaves[leafnum]->match(wld, pla, max_prob, var_wld, sigma_d, oc);
  }
  return flag;
}

// src/mapping/octree.cpp:732-737
bool OctoTree::inside(Vec3& wld)
{
  double hl = quater_length * 2;
  return (wld[0] >= voxel_center[0] - hl && wld[0] <= voxel_center[0] + hl && wld[1] >= voxel_center[1] - hl &&
          wld[1] <= voxel_center[1] + hl && wld[2] >= voxel_center[2] - hl && wld[2] <= voxel_center[2] + hl);
}

// src/mapping/octree.cpp:739-756
void OctoTree::clear_slwd(std::vector<SlideWindow*>& sws)
{
  if (octo_state != 0)
  {
    for (int i = 0; i < 8; i++)
      if (leaves[i] != nullptr) leaves[i]->clear_slwd(sws);
  }
  if (sw != nullptr)
  {
    sw->clear();
    sws.push_back(sw);
    sw = nullptr;
  }
}

// src/mapping/octree.cpp:610-626
void OctoTree::delete_ptr()
{
  for (int i = 0; i < 8; i++)
    if (leaves[i] != nullptr)
    {
      leaves[i]->delete_ptr();
      delete leaves[i];
      leaves[i] = nullptr;
    }
  if (sw != nullptr)
  {
    delete sw;
    sw = nullptr;
  }
}

// ---------------------------------------------------------------------------
// src/mapping/voxel_map.cpp:47-135
void cut_voxel_multi(Globals* G, VoxelMap& feat_map, PVecPtr pvec, int win_count, VoxelMap& feat_tem_map, int wdsize,
                     std::vector<Vec3>& pwld, std::vector<std::vector<SlideWindow*>>& sws)
{
  std::unordered_map<OctoTree*, std::vector<int>> map_pvec;
  int plsize = pvec->size();
  for (int i = 0; i < plsize; i++)
  {
    Vec3& pw = pwld[i];
    VOXEL_LOC position = voxel_key(pw, G->voxel_size);
    auto iter = feat_map.find(position);
    OctoTree* ot = nullptr;
    if (iter != feat_map.end())
    {
      iter->second->isexist = true;
      if (feat_tem_map.find(position) == feat_tem_map.end()) feat_tem_map[position] = iter->second;
      ot = iter->second;
    }
    else
    {
      ot = new OctoTree(G, 0, wdsize);
      ot->voxel_center[0] = (0.5 + position.x) * G->voxel_size;
      ot->voxel_center[1] = (0.5 + position.y) * G->voxel_size;
      ot->voxel_center[2] = (0.5 + position.z) * G->voxel_size;
      ot->quater_length = G->voxel_size / 4.0;
      ot->root_key = position;
      feat_map[position] = ot;
      feat_tem_map[position] = ot;
    }
    map_pvec[ot].push_back(i);
  }

  std::vector<std::pair<OctoTree* const, std::vector<int>>*> octs;
  octs.reserve(map_pvec.size());
  for (auto iter = map_pvec.begin(); iter != map_pvec.end(); iter++) octs.push_back(&(*iter));

  int thd_num = sws.size();
  int g_size = octs.size();
  if (g_size < thd_num) return;
  std::vector<std::thread*> mthreads(thd_num);
  double part = 1.0 * g_size / thd_num;

  int swsize = sws[0].size() / thd_num;
  for (int i = 1; i < thd_num; i++)
  {
    sws[i].insert(sws[i].end(), sws[0].end() - swsize, sws[0].end());
    sws[0].erase(sws[0].end() - swsize, sws[0].end());
  }

  for (int i = 1; i < thd_num; i++)
  {
    mthreads[i] = new std::thread(
        [&](int head, int tail, std::vector<SlideWindow*>& sw) {
          for (int j = head; j < tail; j++)
            for (int k : octs[j]->second) octs[j]->first->allocate(win_count, (*pvec)[k], pwld[k], sw);
        },
        part * i, part * (i + 1), std::ref(sws[i]));
  }
  for (int i = 0; i < thd_num; i++)
  {
    if (i == 0)
    {
      for (int j = 0; j < int(part); j++)
        for (int k : octs[j]->second) octs[j]->first->allocate(win_count, (*pvec)[k], pwld[k], sws[0]);
    }
    else
    {
      mthreads[i]->join();
      delete mthreads[i];
    }
  }
}

// src/mapping/voxel_map.cpp:241-266
int match(Globals* G, VoxelMap& feat_map, Vec3& wld, Plane*& pla, Mat3& var_wld, double& sigma_d, OctoTree*& oc)
{
  int flag = 0;
  VOXEL_LOC position = voxel_key(wld, G->voxel_size);
  auto iter = feat_map.find(position);
  if (iter != feat_map.end())
  {
    double max_prob = 0;
    flag = iter->second->match(wld, pla, max_prob, var_wld, sigma_d, oc);
  }
  return flag;
}

// ---------------------------------------------------------------------------
IMUEKF::IMUEKF()
{
  cov_acc = V3(1, 1, 1);
  cov_gyr = V3(0.01, 0.01, 0.01);
  cov_bias_gyr = V3(1e-4, 1e-4, 1e-4);
  cov_bias_acc = V3(1e-4, 1e-4, 1e-4);
  Lid_rot_to_IMU.setIdentity();
  Lid_offset_to_IMU.setZero();
  last_imu.t = 0;
  for (int i = 0; i < 3; i++) last_imu.gyr[i] = last_imu.acc[i] = 0;
}

// src/estimation/imu_ekf.cpp:13-104 : IMU forward propagation of state + cov
int IMUEKF::propagate(IMUST& xc, std::deque<ImuSample>& imus)
{
  imus.push_front(last_imu);
  if (last_pcl_end_time - pcl_beg_time > 0.01) return -1;  // "LiDAR time regress" (reference exit(0)s)

  imu_poses.clear();
  Vec3 acc_imu = Vec3::Zero(), angvel_avr = Vec3::Zero(), acc_avr, vel_imu(xc.v), pos_imu(xc.p);
  Mat3 R_imu(xc.R);
  Mat15 F_x, cov_w;
  double dt = 0;
  for (size_t it = 0; it + 1 < imus.size(); it++)
  {
    ImuSample& head = imus[it];
    ImuSample& tail = imus[it + 1];
    if (head.t < last_pcl_end_time) continue;

    angvel_avr = V3(0.5 * (head.gyr[0] + tail.gyr[0]), 0.5 * (head.gyr[1] + tail.gyr[1]),
                    0.5 * (head.gyr[2] + tail.gyr[2]));
    acc_avr = V3(0.5 * (head.acc[0] + tail.acc[0]), 0.5 * (head.acc[1] + tail.acc[1]),
                 0.5 * (head.acc[2] + tail.acc[2]));
    angvel_avr -= xc.bg;
    acc_avr = acc_avr * scale_gravity - xc.ba;
    acc_imu = R_imu * acc_avr + xc.g;

    double cur_time = head.t;
    if (cur_time < last_pcl_end_time) cur_time = last_pcl_end_time;
    dt = tail.t - cur_time;
    double offt = cur_time - pcl_beg_time;

    IMUST pose;  // IMUST reused as a tuple (imu_ekf.cpp:62-63): bg := angvel_avr, ba := acc_imu
    pose.t = offt;
    pose.R = R_imu;
    pose.p = pos_imu;
    pose.v = vel_imu;
    pose.bg = angvel_avr;
    pose.ba = acc_imu;
    imu_poses.push_back(pose);

    Mat3 acc_avr_skew = hat(acc_avr);
    Mat3 Exp_f = Exp(angvel_avr, dt);

    F_x.setIdentity();
    cov_w.setZero();
    F_x.setBlock<3, 3>(0, 0, Exp(angvel_avr, -dt));
    F_x.setBlock<3, 3>(0, 9, (-1.0 * Mat3::Identity()) * dt);
    F_x.setBlock<3, 3>(3, 6, Mat3::Identity() * dt);
    F_x.setBlock<3, 3>(6, 0, (-1.0 * R_imu) * acc_avr_skew * dt);
    F_x.setBlock<3, 3>(6, 12, (-1.0 * R_imu) * dt);
    for (int k = 0; k < 3; k++) cov_w(k, k) = cov_gyr[k] * dt * dt;
    Mat3 D = Mat3::Zero();
    for (int k = 0; k < 3; k++) D(k, k) = cov_acc[k];
    cov_w.setBlock<3, 3>(6, 6, R_imu * D * R_imu.transpose() * dt * dt);
    for (int k = 0; k < 3; k++) cov_w(9 + k, 9 + k) = cov_bias_gyr[k] * dt * dt;
    for (int k = 0; k < 3; k++) cov_w(12 + k, 12 + k) = cov_bias_acc[k] * dt * dt;

    xc.cov = F_x * xc.cov * F_x.transpose() + cov_w;

    pos_imu = pos_imu + vel_imu * dt + 0.5 * acc_imu * dt * dt;
    vel_imu = vel_imu + acc_imu * dt;
    R_imu = R_imu * Exp_f;
  }

  double imu_end_time = imus.back().t;
  double note = pcl_end_time > imu_end_time ? 1.0 : -1.0;
  dt = note * (pcl_end_time - imu_end_time);
  xc.v = vel_imu + note * acc_imu * dt;
  xc.R = R_imu * Exp(note * angvel_avr, dt);
  xc.p = pos_imu + note * vel_imu * dt + note * 0.5 * acc_imu * dt * dt;
  xc.t = pcl_end_time;

  // imu_ekf.cpp:95-104: the deque handed on to the pre-integration gets copies of its first / last sample
  // re-stamped to the scan boundaries (integer nanoseconds, rclcpp::Time)
  ImuSample imu1 = imus.front(), imu2 = imus.back();
  imu1.t = (double)static_cast<int64_t>(last_pcl_end_time * 1e9) * 1e-9;
  imu2.t = (double)static_cast<int64_t>(pcl_end_time * 1e9) * 1e-9;
  last_imu = imus.back();
  last_pcl_end_time = pcl_end_time;
  imus.front() = imu1;
  imus.back() = imu2;
  return 0;
}

// src/estimation/imu_ekf.cpp:106-144 : backward per-point deskew
void IMUEKF::deskew(const IMUST& xc, Cloud& pcl_in)
{
  if (point_notime) return;
  if (pcl_in.empty()) return;
  Vec3 acc_imu, angvel_avr, vel_imu, pos_imu;
  Mat3 R_imu;
  long it_pcl = (long)pcl_in.size() - 1;
  for (int i = (int)imu_poses.size() - 1; i >= 0; i--)
  {
    IMUST& head = imu_poses[i];
    R_imu = head.R;
    acc_imu = head.ba;
    vel_imu = head.v;
    pos_imu = head.p;
    angvel_avr = head.bg;
    for (; pcl_in[it_pcl].curvature > head.t; it_pcl--)
    {
      double dt = pcl_in[it_pcl].curvature - head.t;
      Mat3 R_i = R_imu * Exp(angvel_avr, dt);
      Vec3 T_ei = pos_imu + vel_imu * dt + 0.5 * acc_imu * dt * dt - xc.p;
      Vec3 P_i = V3(pcl_in[it_pcl].x, pcl_in[it_pcl].y, pcl_in[it_pcl].z);
      Vec3 P_compensate = Lid_rot_to_IMU.transpose() *
                          (xc.R.transpose() * (R_i * (Lid_rot_to_IMU * P_i + Lid_offset_to_IMU) + T_ei) -
                           Lid_offset_to_IMU);
      pcl_in[it_pcl].x = P_compensate[0];
      pcl_in[it_pcl].y = P_compensate[1];
      pcl_in[it_pcl].z = P_compensate[2];
      if (it_pcl == 0) break;
    }
  }
}

int IMUEKF::motion_blur(IMUST& xc, Cloud& pcl_in, std::deque<ImuSample>& imus)
{
  int r = propagate(xc, imus);
  if (r != 0) return r;
  deskew(xc, pcl_in);
  return 0;
}

// ---------------------------------------------------------------------------
Odom::Odom(const Globals& g) : G(g)
{
  G.mp.resize(G.win_size);
  for (int i = 0; i < G.win_size; i++) G.mp[i] = i;  // node.cpp:431-435
  sws.resize(G.thread_num);                          // node.cpp:289
}

Odom::~Odom()
{
  for (auto& kv : surf_map)
  {
    kv.second->delete_ptr();
    delete kv.second;
  }
  for (auto& v : sws)
    for (SlideWindow* s : v) delete s;
  for (IMU_PRE* f : imu_pre_buf) delete f;
}

// src/pipeline/odometry.cpp:64-255 (use_vnc == false path; VNC terms are additive, :151-190)
bool Odom::LioStateEstimation(PVecPtr pptr, int max_iter_override)
{
  IMUST x_prop = x_curr;
  const int num_max_iter = max_iter_override > 0 ? max_iter_override : 20;
  bool EKF_stop_flg = false, flg_EKF_converged = false;
  Mat15 G_, H_T_H, I_STATE;
  G_.setZero();
  H_T_H.setZero();
  I_STATE.setIdentity();
  int rematch_num = 0;
  int match_num = 0;
  int psize = pptr->size();
  std::vector<OctoTree*> octos(psize, nullptr);
  Mat3 nnt;
  nnt.setZero();
  Mat15 cov_inv = inverse(x_curr.cov);
  iter_dumps.clear();
  last_iters = 0;

  for (int iterCount = 0; iterCount < num_max_iter; iterCount++)
  {
    Mat6 HTH;
    HTH.setZero();
    Vec6 HTz;
    HTz.setZero();
    Mat3 rot_var = x_curr.cov.block<3, 3>(0, 0);
    Mat3 tsl_var = x_curr.cov.block<3, 3>(3, 3);
    match_num = 0;
    nnt.setZero();
    IekfIterDump* dump = nullptr;
    if (dump_iters)
    {
      iter_dumps.emplace_back();
      dump = &iter_dumps.back();
      dump->keys.resize(3 * (size_t)psize);
      dump->codes.assign(psize, -1);
      dump->flags.assign(psize, 0);
      dump->sigma.assign(psize, 0.0);
      memcpy(dump->R, x_curr.R.d, sizeof(dump->R));
      memcpy(dump->p, x_curr.p.d, sizeof(dump->p));
    }

    for (int i = 0; i < psize; i++)
    {
      pointVar& pv = pptr->at(i);
      Mat3 phat = hat(pv.pnt);
      Mat3 var_world = x_curr.R * pv.var * x_curr.R.transpose() + phat * rot_var * phat.transpose() + tsl_var;
      Vec3 wld = x_curr.R * pv.pnt + x_curr.p;

      double sigma_d = 0;
      Plane* pla = nullptr;
      int flag = 0;
      if (octos[i] != nullptr && octos[i]->inside(wld))
      {
        double max_prob = 0;
        flag = octos[i]->match(wld, pla, max_prob, var_world, sigma_d, octos[i]);
      }
      else
      {
        flag = match(&G, surf_map, wld, pla, var_world, sigma_d, octos[i]);
      }
      if (dump)
      {
        VOXEL_LOC k = voxel_key(wld, G.voxel_size);
        dump->keys[3 * (size_t)i + 0] = k.x;
        dump->keys[3 * (size_t)i + 1] = k.y;
        dump->keys[3 * (size_t)i + 2] = k.z;
        dump->flags[i] = flag ? 1 : 0;
        if (flag)
        {
          dump->codes[i] = octos[i]->code();
          dump->sigma[i] = sigma_d;
        }
      }

      if (flag)
      {
        Plane& pp = *pla;
        double R_inv = 1.0 / (0.0005 + sigma_d);
        double resi = dot(pp.normal, wld - pp.center);
        Vec6 jac;
        Vec3 h = phat * x_curr.R.transpose() * pp.normal;
        for (int k = 0; k < 3; k++)
        {
          jac[k] = h[k];
          jac[3 + k] = pp.normal[k];
        }
        for (int c = 0; c < 6; c++)
          for (int r = 0; r < 6; r++) HTH(r, c) = HTH(r, c) + (R_inv * jac[r]) * jac[c];
        for (int r = 0; r < 6; r++) HTz[r] = HTz[r] - (R_inv * jac[r]) * resi;
        for (int c = 0; c < 3; c++)
          for (int r = 0; r < 3; r++) nnt(r, c) = nnt(r, c) + pp.normal[r] * pp.normal[c];
        match_num++;
      }
    }
    if (dump)
    {
      memcpy(dump->HTH, HTH.d, sizeof(dump->HTH));
      memcpy(dump->HTz, HTz.d, sizeof(dump->HTz));
      memcpy(dump->nnt, nnt.d, sizeof(dump->nnt));
      dump->match_num = match_num;
    }
    last_iters = iterCount + 1;

    H_T_H.setBlock<6, 6>(0, 0, HTH);
    Mat15 K_1 = inverse(H_T_H + cov_inv);
    Mat<15, 6> K6 = K_1.block<15, 6>(0, 0);
    Mat<15, 6> G6 = K6 * HTH;
    G_.setBlock<15, 6>(0, 0, G6);
    Vec15 vec = x_prop - x_curr;
    Vec15 solution = K6 * HTz + vec - G6 * vec.block<6, 1>(0, 0);
    x_curr += solution;

    Vec3 rot_add = solution.block<3, 1>(0, 0);
    Vec3 tra_add = solution.block<3, 1>(3, 0);
    EKF_stop_flg = false;
    flg_EKF_converged = false;
    if ((norm(rot_add) * 57.3 < 0.01) && (norm(tra_add) * 100 < 0.015)) flg_EKF_converged = true;
    if (flg_EKF_converged || ((rematch_num == 0) && (iterCount == num_max_iter - 2))) rematch_num++;
    if (rematch_num >= 2 || (iterCount == num_max_iter - 1))
    {
      x_curr.cov = (I_STATE - G_) * x_curr.cov;
      EKF_stop_flg = true;
    }
    if (EKF_stop_flg) break;
  }

  SelfAdjointEigen3 saes(nnt);
  return !(saes.values[0] < 14);
}

// src/pipeline/local_mapping.cpp:144-201 (production overload; the BA factor
// containers reduce to marking opt_state, see OctoTree::margi above)
void Odom::multi_recut(VoxelMap& feat_map, int win_count, std::vector<IMUST>& xs,
                       std::vector<std::vector<SlideWindow*>>& sws)
{
  int thd_num = G.thread_num;
  std::vector<std::vector<OctoTree*>> octss(thd_num);
  int g_size = feat_map.size();
  if (g_size < thd_num) return;
  std::vector<std::thread*> mthreads(thd_num);
  double part = 1.0 * g_size / thd_num;
  int cnt = 0;
  for (auto iter = feat_map.begin(); iter != feat_map.end(); iter++)
  {
    octss[cnt].push_back(iter->second);
    if (octss[cnt].size() >= part && cnt < thd_num - 1) cnt++;
  }
  auto recut_func = [](int win_count, std::vector<OctoTree*>& oct, std::vector<IMUST> xxs,
                       std::vector<SlideWindow*>& sw) {
    for (OctoTree* oc : oct) oc->recut(win_count, xxs, sw);
  };
  for (int i = 1; i < thd_num; i++)
    mthreads[i] = new std::thread(recut_func, win_count, std::ref(octss[i]), xs, std::ref(sws[i]));
  for (int i = 0; i < thd_num; i++)
  {
    if (i == 0)
      recut_func(win_count, octss[i], xs, sws[i]);
    else
    {
      mthreads[i]->join();
      delete mthreads[i];
    }
  }
  for (size_t i = 1; i < sws.size(); i++)
  {
    sws[0].insert(sws[0].end(), sws[i].begin(), sws[i].end());
    sws[i].clear();
  }
  int counter = 0;
  for (auto iter = feat_map.begin(); iter != feat_map.end(); iter++) tras_opt_mark(iter->second, counter);
}

// src/pipeline/local_mapping.cpp:17-84
void Odom::multi_margi(VoxelMap& feat_map, int win_count, std::vector<IMUST>& xs, std::vector<SlideWindow*>& sw)
{
  int thd_num = G.thread_num;
  std::vector<std::vector<OctoTree*>> octs(thd_num);
  int g_size = feat_map.size();
  if (g_size < thd_num) return;
  std::vector<std::thread*> mthreads(thd_num);
  double part = 1.0 * g_size / thd_num;
  int cnt = 0;
  for (auto iter = feat_map.begin(); iter != feat_map.end(); iter++)
  {
    iter->second->jour = jour;  // local_mapping.cpp:36
    octs[cnt].push_back(iter->second);
    if (octs[cnt].size() >= part && cnt < thd_num - 1) cnt++;
  }
  auto margi_func = [](int win_cnt, std::vector<OctoTree*>* oct, std::vector<IMUST> xxs) {
    for (OctoTree* oc : *oct) oc->margi(win_cnt, 1, xxs);
  };
  for (int i = 1; i < thd_num; i++) mthreads[i] = new std::thread(margi_func, win_count, &octs[i], xs);
  for (int i = 0; i < thd_num; i++)
  {
    if (i == 0)
      margi_func(win_count, &octs[i], xs);
    else
    {
      mthreads[i]->join();
      delete mthreads[i];
    }
  }
  for (auto iter = feat_map.begin(); iter != feat_map.end();)
  {
    if (iter->second->isexist)
      iter++;
    else
    {
      iter->second->clear_slwd(sw);
      feat_map.erase(iter++);
    }
  }
}

// src/pipeline/local_mapping.cpp:434-451 and 489-546
void Odom::map_update(PVecPtr pptr, std::deque<ImuSample>* imus)
{
  const int mgsize = 1;
  win_count++;
  x_buf.push_back(x_curr);
  pvec_buf.push_back(pptr);
  if (win_count > 1)
  {
    IMU_PRE* f = nullptr;
    if (imus)
    {
      f = new IMU_PRE(x_buf[win_count - 2].bg, x_buf[win_count - 2].ba);
      f->push_imu(*imus, ba_noise);
    }
    imu_pre_buf.push_back(f);
  }

  double t1 = now_s();
  cut_voxel_multi(&G, surf_map, pvec_buf[win_count - 1], win_count - 1, surf_map_slide, G.win_size, pwld, sws);
  double t2 = now_s();
  multi_recut(surf_map_slide, win_count, x_buf, sws);
  double t3 = now_s();
  t_insert = t2 - t1;
  t_recut = t3 - t2;
  if (ba_probe && win_count >= G.win_size)
  {
    ba_factors.clear();
    ba_factors.win_size = G.win_size;
    if ((int)surf_map_slide.size() >= G.thread_num)  // multi_recut's early-out (local_mapping.cpp:150-154)
      for (auto iter = surf_map_slide.begin(); iter != surf_map_slide.end(); iter++)
        tras_opt_collect(iter->second, ba_factors);
    ba_xs = x_buf;
  }
  t_margi = 0;
  window_tail();
}

// src/pipeline/local_mapping.cpp:489-546
void Odom::window_tail(LidarFactor* vh, std::vector<OctoTree*>* vh_nodes)
{
  const int mgsize = 1;
  if (win_count >= G.win_size)
  {
    bool all_imu = (int)imu_pre_buf.size() == win_count - 1;
    for (IMU_PRE* f : imu_pre_buf) all_imu = all_imu && f != nullptr;
    LidarFactor voxhess;
    std::vector<OctoTree*> nodes;
    if (vh == nullptr && if_BA && all_imu)
    {
      // the factors tras_opt collected in multi_recut (none when multi_recut took its early-out,
      // local_mapping.cpp:150-154)
      voxhess.win_size = G.win_size;
      if ((int)surf_map_slide.size() >= G.thread_num)
        for (auto iter = surf_map_slide.begin(); iter != surf_map_slide.end(); iter++)
          tras_opt_collect(iter->second, voxhess, &nodes);
      vh = &voxhess;
      vh_nodes = &nodes;
    }
    if (if_BA && all_imu)
    {
      // local_mapping.cpp:492-497: LI_BA_Optimizer
      ba_last_iters = ba_damping_iter(x_buf, *vh, imu_pre_buf, imu_coef, nullptr);
      ba_runs++;
    }
    if (vh != nullptr)
    {
      // OctoTree::margi takes the factors' (possibly re-evaluated) pcr_add / eig back (octree.cpp:410-416)
      for (size_t a = 0; a < vh_nodes->size(); a++)
      {
        (*vh_nodes)[a]->pcr_add = vh->pcr_adds[a];
        (*vh_nodes)[a]->eig_value = vh->eig_values[a];
        (*vh_nodes)[a]->eig_vector = vh->eig_vectors[a];
      }
    }
    x_curr.R = x_buf[win_count - 1].R;
    x_curr.p = x_buf[win_count - 1].p;
    double t5 = now_s();
    multi_margi(surf_map_slide, win_count, x_buf, sws[0]);
    t_margi = now_s() - t5;

    if ((win_base + win_count) % 10 == 0)  // local_mapping.cpp:509-519
    {
      double spat = norm(x_curr.p - last_pos);
      if (spat > 0.5)
      {
        jour += spat;
        last_pos = x_curr.p;
        release_flag = true;
      }
    }

    for (int i = 0; i < G.win_size; i++)
    {
      G.mp[i] += mgsize;
      if (G.mp[i] >= G.win_size) G.mp[i] -= G.win_size;
    }
    for (int i = mgsize; i < win_count; i++)
    {
      x_buf[i - mgsize] = x_buf[i];
      PVecPtr pvec_tem = pvec_buf[i - mgsize];
      pvec_buf[i - mgsize] = pvec_buf[i];
      pvec_buf[i] = pvec_tem;
    }
    for (int i = win_count - mgsize; i < win_count; i++)
    {
      x_buf.pop_back();
      pvec_buf.pop_back();
      delete imu_pre_buf.front();
      imu_pre_buf.pop_front();
    }
    win_base += mgsize;
    win_count -= mgsize;
  }
}

// src/mapping/octree.cpp:597-608
void Odom::tras_ptr(OctoTree* ot, std::vector<OctoTree*>& octos_release)
{
  if (ot->octo_state == 1)
    for (int i = 0; i < 8; i++)
      if (ot->leaves[i] != nullptr)
      {
        octos_release.push_back(ot->leaves[i]);
        tras_ptr(ot->leaves[i], octos_release);
      }
}

// src/pipeline/local_mapping.cpp:317-341
int Odom::idle_release(int horizon, int* nodes_freed)
{
  if (nodes_freed) *nodes_freed = 0;
  if (!release_flag) return 0;
  release_flag = false;
  std::vector<OctoTree*> octos;
  int roots = 0;
  for (auto iter = surf_map.begin(); iter != surf_map.end();)
  {
    int dis = jour - iter->second->jour;  // double -> int, truncation toward zero
    if (dis < horizon || surf_map_slide.count(iter->first))  // (see below for the second condition)
      iter++;
    else
    {
      octos.push_back(iter->second);
      tras_ptr(iter->second, octos);
      surf_map.erase(iter++);
      roots++;
    }
  }
  if (nodes_freed) *nodes_freed = (int)octos.size();
  // `delete octos[i]` (the reference's OctoTree has no destructor: the children are in the list themselves). A
  // root that is still in surf_map_slide would dangle there - with the reference's 700 m horizon it cannot be
  // (its stamp is at most one marginalisation old), with the small horizons the tests use it could: such roots
  // are kept, here, in the reference harness and on the device alike.
  for (OctoTree* o : octos) delete o;
  return roots;
}

// src/pipeline/local_mapping.cpp:389-546
int Odom::step(Cloud& pcl_curr, double pcl_beg_time, std::deque<ImuSample>& imus, bool iekf_on_full, int max_iter)
{
  double t0 = now_s();
  odom_ekf.pcl_beg_time = pcl_beg_time;
  odom_ekf.pcl_end_time = pcl_beg_time + pcl_curr.back().curvature;  // sync.cpp:40
  if (odom_ekf.motion_blur(x_curr, pcl_curr, imus) != 0) return -1;

  Cloud pl_down = pcl_curr;
  down_sampling_voxel(pl_down, G.down_size);
  if (pl_down.size() < 2000)
  {
    pl_down = pcl_curr;
    down_sampling_voxel(pl_down, G.down_size / 2);
  }
  last_down = pl_down;

  PVecPtr pptr(new PVec);
  var_init(extrin_para, pl_down, pptr, G.dept_err, G.beam_err);
  PVecPtr no_ds_pptr(new PVec);
  if (iekf_on_full)
  {
    Cloud pcl_curr_temp = pcl_curr;
    var_init(extrin_para, pcl_curr_temp, no_ds_pptr, G.dept_err, G.beam_err);
  }
  last_pptr = pptr;
  last_full_pptr = no_ds_pptr;

  if (LioStateEstimation(iekf_on_full ? no_ds_pptr : pptr, max_iter))
  {
    if (degrade_cnt > 0) degrade_cnt--;
  }
  else
    degrade_cnt++;

  pwld.clear();
  pvec_update(pptr, x_curr, pwld);
  t_odom = now_s() - t0;

  map_update(pptr, &imus);
  return 0;
}

void Odom::bootstrap(Cloud& pcl_deskewed, const IMUST& x_known)
{
  x_curr = x_known;
  Cloud pl_down = pcl_deskewed;
  down_sampling_voxel(pl_down, G.down_size);
  if (pl_down.size() < 2000)
  {
    pl_down = pcl_deskewed;
    down_sampling_voxel(pl_down, G.down_size / 2);
  }
  last_down = pl_down;
  PVecPtr pptr(new PVec);
  var_init(extrin_para, pl_down, pptr, G.dept_err, G.beam_err);
  last_pptr = pptr;
  pwld.clear();
  pvec_update(pptr, x_curr, pwld);
  map_update(pptr);
}
}  // namespace vo
