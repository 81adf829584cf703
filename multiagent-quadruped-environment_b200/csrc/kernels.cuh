// kernels.cuh -- host-side launcher prototypes shared between the kernel translation units and api.cu.
#pragma once
#include "common.cuh"

struct PolicyWeightsDev {           // device copies laid out for the kernels (api.cu builds them)
    const float *w0cat;             // [768][30][80]  rows 0..255 adapt.0, 256..767 body.0[:, :2100]; age blocks padded 70 -> 80
    const float *b0cat;             // [768]
    const float *wlat;              // [512][2]      body.0[:, 2100:2102]  (the latent columns of cat(h, latent))
    const float *aw1, *ab1, *aw2, *ab2;                         // adapt [128][256], [2][128]
    const float *bw1, *bb1, *bw2, *bb2, *bw3, *bb3;             // body  [256][512], [128][256], [12][128]
};
struct PolicyTcWeights {            // bf16 hi/lo planes, pre-tiled (policy_tc.cu)
    void *blob; size_t bytes; int rows;
    const void *l0_hi, *l0_lo;       // layer 0 weights
    const void *t_hi[5], *t_lo[5];   // tail layer weights
    void *p_hi[5], *p_lo[5];         // activation planes between layers (sized for `rows`)
};
struct PolicyScratch { float *Z, *T1, *T2, *T3, *latent, *act, *Zold; };   // Zold: layer-0 partial sums of the 29 known frames (incremental layer 0)

extern "C" {
cudaError_t mqe_launch_substeps(const DevParams &p, int nsub, int maxpair, cudaStream_t st);
cudaError_t mqe_launch_balance_tasks(const DevParams &p, int *order, int spread, int maxpair, cudaStream_t st);
cudaError_t mqe_launch_post(const DevParams &p, unsigned int step_count, cudaStream_t st);
cudaError_t mqe_launch_reset_all(const DevParams &p, cudaStream_t st);
cudaError_t mqe_launch_set_root_indexed(const DevParams &p, const float *src, const int *ids, int n, cudaStream_t st);
cudaError_t mqe_launch_set_dof_indexed(const DevParams &p, const float *src, const int *ids, int n, cudaStream_t st);
cudaError_t mqe_launch_policy_frame(const DevParams &p, const float *d_actions, int head, cudaStream_t st);
cudaError_t mqe_launch_policy_finish(const DevParams &p, const float *act, cudaStream_t st);
cudaError_t mqe_launch_policy_l0_fp32(const PolicyWeightsDev &w, const PolicyScratch &s, const float *ring, int head, int M, cudaStream_t st);
cudaError_t mqe_launch_policy_tail(const PolicyWeightsDev &w, const PolicyScratch &s, int M, cudaStream_t st, int *launches);
cudaError_t mqe_launch_history_to_ring(const float *hist, float *ring, unsigned short *hi, unsigned short *lo, int rows, cudaStream_t st);
cudaError_t mqe_launch_actuator(const float *act_w, const float *x, int rows, float *out, cudaStream_t st);
// tensor-core policy layer 0 (policy_tc.cu)
int mqe_policy_tc_prepare(const MqeWeights *w, int rows, PolicyTcWeights *out, cudaStream_t st);
cudaError_t mqe_launch_policy_l0_tc(const PolicyTcWeights &w, const float *b0cat, const unsigned short *hist_hi, const unsigned short *hist_lo,
                                    int head, int rows, int passes, float *Z, int planes_out, const int *ctr, cudaStream_t st);
cudaError_t mqe_launch_policy_tail_tc(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, int M, int passes,
                                      cudaStream_t st, int *launches);
cudaError_t mqe_launch_policy_tc_fused(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const DevParams &p,
                                       const unsigned short *hist_hi, const unsigned short *hist_lo, int head, int M, int passes, const int *ctr,
                                       int finish, cudaStream_t st, int *launches);
cudaError_t mqe_launch_policy_l0_old(const PolicyTcWeights &w, const unsigned short *hist_hi, const unsigned short *hist_lo, int head_next,
                                     int M, int passes, float *Zold, const int *ctr, int mtile0, int mtiles, cudaStream_t st);
cudaError_t mqe_launch_policy_tc_incremental(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const DevParams &p,
                                             const unsigned short *hist_hi, const unsigned short *hist_lo, int head, int M, int passes,
                                             const int *ctr, int finish, cudaStream_t st, int *launches,
                                             int early_tiles, int head_next, cudaStream_t aux, cudaEvent_t ev, cudaEvent_t join_before_tail);
cudaError_t mqe_launch_task_gather(const DevParams &p, const WrapParams &w, int mode, cudaStream_t st);
cudaError_t mqe_launch_policy_tc_forked(const PolicyTcWeights &w, const PolicyWeightsDev &pw, const PolicyScratch &s, const unsigned short *hist_hi,
                                        const unsigned short *hist_lo, int head, int M, int passes, const int *ctr, cudaStream_t st, cudaStream_t aux,
                                        cudaEvent_t ev_fork, cudaEvent_t ev_join, int *launches);
cudaError_t mqe_launch_joint_actions(const DevParams &p, const float *joint_actions, cudaStream_t st);
cudaError_t mqe_launch_gather_exchange(const DevParams &p, const GatherParams &g, cudaStream_t st);
cudaError_t mqe_substeps_configure(const DevParams &p, int maxpair);
size_t mqe_substeps_smem_bytes(int N, int A, int Pd, int E, int maxpair);
size_t mqe_substeps_row_scratch_floats(int N, int A);
size_t mqe_substeps_prow_scratch_floats(int N, int maxpair);
size_t mqe_substeps_pdesc_scratch_floats(int N, int maxpair);
}
