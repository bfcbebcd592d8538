// pixel_launch.inl — host launcher body shared by pixel_f64.cu / pixel_f32.cu (REAL, LAUNCH_NAME set
// by the including TU).
int LAUNCH_NAME(const DevSceneT<REAL>& sc, const DevFrameT<REAL>& fr, const RowMap& rm, uint8_t* d_out,
                unsigned long long* d_counters, cudaStream_t st, uint32_t* launches) {
  const long long total = (long long)rm.n_local() * fr.width;
  if (total <= 0) return RTB_OK;
  const int threads = 128;
  size_t smem = (size_t)sc.n_spheres * 4 * sizeof(REAL);
  int in_smem = 1;
  if (smem > g_rtb.smem_optin || smem == 0) { smem = 0; in_smem = 0; }
  if (smem > 48 * 1024)
    RTB_CUDA(cudaFuncSetAttribute(pixel_kernel<REAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // Trees differ wildly in size (81 % of demo pixels are one node, 2.4 % reach the depth cap), so
  // many small CTAs balance better than one persistent wave: 128-thread CTAs, at most 64 per SM
  // worth of grid, grid-stride beyond that.
  long long blocks = (total + threads - 1) / threads;
  const long long max_blocks = (long long)g_rtb.sm_count * 64;
  if (blocks > max_blocks) blocks = max_blocks;
  pixel_kernel<REAL><<<(unsigned)blocks, threads, smem, st>>>(sc, fr, rm, d_out, d_counters, in_smem);
  RTB_CUDA(cudaGetLastError());
  if (launches) (*launches)++;
  return RTB_OK;
}
