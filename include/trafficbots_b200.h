/*
 * trafficbots_b200 -- C ABI of the B200-native TrafficBots hot path (scene encoding + closed-loop rollout).
 *
 * The reference (zhejz/TrafficBots) is pure Python/PyTorch and has NO C/FFI boundary; its "plugin API" for this
 * path is a set of nn.Module methods selected through Hydra `_target_`s.  Each entry point below
 * names the reference method(s) it replaces (paths relative to the reference's `src/`); the Python mirror of
 * those methods (trafficbots_b200/pl_modules, trafficbots_b200/models) binds this library with ctypes, and
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless it says "host"; all tensors are dense, row-major, in the
 *     reference's batch schema (`data_modules/data_h5_womd.py:85-173`); `bool` tensors are 1 byte per element.
 *   - the caller owns and allocates every buffer; the library keeps no global mutable state that affects results (a launch
 *     counter for diagnostics, `tb_launch_count`, is the only process-wide variable) and never
 *     allocates device memory; every call is asynchronous on the `stream` argument (a `cudaStream_t`).
 *   - return value: TB_OK or a negative tb_status; nothing is thrown across the ABI.
 *   - inputs, outputs and accumulation are fp32 (the reference's eval-mode dtype); contractions on the tensor pipe use
 *     bf16x3 split operands (x = hi + lo, three bf16 MMAs per product, fp32 accumulate in tensor memory): ~2^-17 relative
 *     error per product; the 90-step closed loop stays within 2e-3 m of the fp32 reference (tests/test_gpu_parity.py).
 */
#ifndef TRAFFICBOTS_B200_H_
#define TRAFFICBOTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum tb_status {
  TB_OK = 0,
  TB_ERR_BAD_SHAPE = -1,   /* a dimension is out of the supported range */
  TB_ERR_NULL = -2,        /* a required pointer is NULL */
  TB_ERR_LAUNCH = -3,      /* a CUDA launch failed (cudaGetLastError) */
  TB_ERR_UNSUPPORTED = -4, /* configuration outside the default `configs/model/traffic_bots.yaml` surface */
  TB_ERR_ALIGN = -5        /* a float buffer is not 16-byte aligned */
} tb_status;

/* Fixed by the default model config (configs/model/traffic_bots.yaml): hidden_dim 128, n_head 4,
 * d_feedforward 128, pe_dim 96, 3 GRU layers, latent_dim 16; data: 20 nodes per polyline, 11 polyline types,
 * 5 traffic-light states. */
#define TB_HIDDEN 128
#define TB_LATENT 16
#define TB_PL_NODE 20
#define TB_PL_TYPE 11
#define TB_TL_STATE 5
#define TB_N_OPT_VIOLATION 8 /* tb_rule_checks */
#define TB_N_VIOLATION 6 /* outside_map, outside_map_this_step, goal_reached, goal_reached_this_step,
                            dest_reached, dest_reached_this_step (utils/traffic_rule_checker.py:499-515) */

typedef struct TbDims {
  int32_t n_scene;     /* S: scenes in the batch */
  int32_t n_mode;      /* K: joint futures per scene; scene-mode b = s*K + k (waymo_motion.py:487-548) */
  int32_t n_agent;     /* A */
  int32_t n_pl;        /* P: map polylines */
  int32_t n_tl;        /* TL: traffic-light stop points */
  int32_t n_step_hist; /* history frames = time_step_current + 1 (11); also the number of TL frames */
  int32_t n_step_gt;   /* frames in the GT tensors used for overriding: 91 (train/val) or 11 (test) */
  int32_t n_step;      /* decode steps stored in the outputs = time_step_end (90) */
  int32_t n_cta_per_mode; /* CTAs (one thread-block cluster) per scene-mode in the persistent decode kernel: 1, 2 or 4;
                             0 = chosen from the batch size so that one launch fills the GPU.  Callers that keep several
                             batches in flight on separate streams pass 1 (no redundant work between cluster ranks).
                             tb_rollout_state_bytes depends on it: use the same value for every call on one state buffer. */
} TbDims;

/* ---------------------------------------------------------------- parameters ------------------------- */

/* Number of tensors in the packed parameter blob, and for tensor i its reference state_dict key, shape
 * (cols == 0 for 1-D tensors) -- the order in which tb_pack_weights expects its pointers. */
int32_t tb_weight_count(void);
const char* tb_weight_name(int32_t i);
int32_t tb_weight_rows(int32_t i);
int32_t tb_weight_cols(int32_t i);
size_t tb_packed_weight_bytes(void);

/* Re-lays the reference's fp32 parameters (host array of `tb_weight_count()` device pointers, each a
 * contiguous tensor exactly as in `WaymoMotion.state_dict()`) into the kernel layout.  Replaces
 * `load_state_dict` (run.py:40-44).  `packed`: tb_packed_weight_bytes() bytes. */
int32_t tb_pack_weights(const float* const* params_host_array, float* packed, void* stream);

/* ---------------------------------------------------------------- scene encoding ---------------------- */

typedef struct TbSceneIn { /* raw batch tensors (what SceneCentricInput consumes, data_modules/sc_input.py:98-140) */
  const uint8_t* map_valid;      /* [S,P,20]      map/valid   */
  const uint8_t* map_type;       /* [S,P,11]      map/type    */
  const float* map_pos;          /* [S,P,20,2]    map/pos     */
  const float* map_dir;          /* [S,P,20,2]    map/dir     */
  const uint8_t* agent_valid;    /* [S,Th,A]      history/agent/valid */
  const float* agent_pos;        /* [S,Th,A,2]    */
  const float* agent_yaw;        /* [S,Th,A,1]    history/agent/yaw_bbox */
  const float* agent_vel;        /* [S,Th,A,2]    */
  const float* agent_spd;        /* [S,Th,A,1]    */
  const float* agent_yaw_rate;   /* [S,Th,A,1]    */
  const float* agent_acc;        /* [S,Th,A,1]    */
  const float* agent_size;       /* [S,A,3]       */
  const uint8_t* agent_type;     /* [S,A,3]       */
  const uint8_t* tl_valid;       /* [S,Th,TL]     history/tl_stop/valid */
  const uint8_t* tl_state;       /* [S,Th,TL,5]   */
  const float* tl_pos;           /* [S,Th,TL,2]   */
  const float* tl_dir;           /* [S,Th,TL,2]   */
} TbSceneIn;

typedef struct TbSceneOut { /* == the dict returned by TrafficBots.encode_input_features (traffic_bots.py:146-151) */
  float* map_feature;            /* [S,P,128]     */
  uint8_t* map_feature_valid;    /* [S,P]         */
  float* agent_feature;          /* [S,Th,A,128]  */
  float* tl_feature;             /* [S,Th,TL,128] */
  /* loop-invariant attention operands, projected once per scene (the reference re-projects them at every
   * decode step and layer, attention.py:86): K|V = LN_tgt(tgt) W_kv + b for the 3 agent->map layers and the
   * 3 agent->traffic-light layers of the policy (traffic_bots.py:205-219). */
  float* kv_map;                 /* [3,S,P,256]      */
  float* kv_tl;                  /* [3,S,Th,TL,256]  */
  /* the same K|V as tensor-core operand blocks (see tb_kv_tc_bytes): only the VALID keys, compacted, in 64-key tiles of
   * 64 KB = bf16 [K hi | K lo | V^T hi | V^T lo] in the 128-byte-swizzled K-major UMMA layout, plus the key counts. */
  uint8_t* kv_map_tc;            /* [3,S,ceil(P/64)] x 64 KB      */
  uint8_t* kv_tl_tc;             /* [3,S,Th,ceil(TL/64)] x 64 KB  */
  int32_t* n_key_map;            /* [S]     valid polylines per scene          */
  int32_t* n_key_tl;             /* [S,Th]  valid traffic lights per scene-frame */
} TbSceneOut;

/* bytes of kv_map_tc (which = 0) / kv_tl_tc (which = 1) for these dims */
size_t tb_kv_tc_bytes(const TbDims* dims, int32_t which);

/* scratch needed by tb_encode_scene */
size_t tb_encode_workspace_bytes(const TbDims* dims);

/* SceneCentricInput.forward + TrafficBots.encode_input_features (data_modules/sc_input.py:98-140,
 * models/traffic_bots.py:109-151, models/modules/map_encoder.py:58-115, input_pe_encoder.py:41-61). */
int32_t tb_encode_scene(const TbDims* dims, const TbSceneIn* in, const float* packed, const TbSceneOut* out,
                        void* workspace, void* stream);
/* With `in->map_valid == NULL` only the agent / traffic-light halves run (map outputs untouched): the reference encodes the
 * same map up to three times per step on aliased inputs (`input/`, `latent_prior/`, `latent_post/`; sc_latent.py:143-148,
 * waymo_motion.py:581-583); callers encode it once and re-use map_feature / kv_map for the posterior pass. */

/* ---------------------------------------------------------------- building blocks --------------------- */

/* Identifies one TransformerBlock of the reference (models/modules/transformer.py:53-95). */
typedef enum tb_block {
  TB_BLOCK_MAP_DENSETNT = 0,   /* model.map_encoder.transformer_densetnt (3 layers)   */
  TB_BLOCK_MAP_SELF_ATTN = 1,  /* model.map_encoder.transformer_self_attn (1 layer)   */
  TB_BLOCK_AS2PL = 2,          /* model.transformer_as2pl (3)                          */
  TB_BLOCK_AS2TL = 3,          /* model.transformer_as2tl (3)                          */
  TB_BLOCK_INTERACTION = 4,    /* model.agent_interaction.transformer (3)              */
  TB_BLOCK_LATENT_PRIOR_INT = 5, /* model.latent_encoder.agent_interaction_prior.transformer (3) */
  TB_BLOCK_LATENT_POST_INT = 6   /* model.latent_encoder.agent_interaction_post.transformer (3)  */
} tb_block;

/* K|V projection of one layer: kv[row] = LN_tgt(tgt[row]) W_kv^T + b_kv  (transformer.py:192 + attention.py:86).
 * tgt [n_row,128] -> kv [n_row,256]. */
int32_t tb_kv_project(int32_t block, int32_t layer, const float* tgt, int64_t n_row, const float* packed,
                      float* kv, void* stream);

/* One pre-LN cross-attention layer, TransformerCrossAttention.forward (transformer.py:186-237) with
 * Attention.forward (attention.py:79-146), K|V given by tb_kv_project.
 *   src [n_batch,n_src,128], src_valid [n_batch,n_src]; the keys of batch element b are
 *   kv[(b / kv_share) ...]: kv [n_batch/kv_share, n_key, 256], key_valid [n_batch/kv_share, n_key];
 *   mask_self != 0 additionally disables key j for query j (the eye attn_mask of agent_interaction.py:57-59).
 * Rows without any enabled key get a zero attention output (attention.py:101-107,144-146); rows with
 * src_valid == 0 are zeroed at the end (transformer.py:236-237).  dst may alias src. */
int32_t tb_xlayer(int32_t block, int32_t layer, const float* src, const uint8_t* src_valid, int32_t n_batch,
                  int32_t n_src, const float* kv, const uint8_t* key_valid, int32_t n_key, int32_t kv_share,
                  int32_t mask_self, const float* packed, float* dst, void* stream);

/* The same layer with the keys given as compacted tensor-core key blocks (TbSceneOut.kv_map_tc / kv_tl_tc layout: 64 keys per
 * 64 KB block, `n_key[i]` valid keys in set i, ceil(n_key_max / 64) blocks per set) and evaluated on the tensor pipe (bf16x3).
 * No self-mask (compaction drops the key index).  Blocks: TB_BLOCK_MAP_SELF_ATTN, TB_BLOCK_AS2PL, TB_BLOCK_AS2TL,
 * TB_BLOCK_INTERACTION.  Used for the map self-attention inside tb_encode_scene and for the latent encoder. */
int32_t tb_xlayer_tc(int32_t block, int32_t layer, const float* src, const uint8_t* src_valid, int32_t n_batch, int32_t n_src,
                     const uint8_t* key_blocks, const int32_t* n_key, int32_t n_key_max, int32_t kv_share, const float* packed,
                     float* dst, void* stream);

/* ---------------------------------------------------------------- closed-loop rollout ----------------- */

typedef struct TbRolloutIn {
  /* per scene (shared by the K modes of the scene) */
  const float* map_feature;        /* [S,P,128]   TbSceneOut.map_feature (destination feature gather) */
  const uint8_t* map_feature_valid;/* [S,P]       */
  const float* kv_map;             /* [3,S,P,256] */
  const float* kv_tl;              /* [3,S,Th,TL,256] */
  const uint8_t* tl_valid;         /* [S,Th,TL]   */
  const uint8_t* gt_valid;         /* [S,Tg,A]    agent/valid   (Tg = n_step_gt) */
  const float* gt_pos;             /* [S,Tg,A,2]  agent/pos      */
  const float* gt_yaw;             /* [S,Tg,A,1]  agent/yaw_bbox */
  const float* gt_spd;             /* [S,Tg,A,1]  agent/spd      */
  const float* gt_vel;             /* [S,Tg,A,2]  agent/vel      */
  const float* gt_acc;             /* [S,Tg,A,1]  agent/acc      */
  const float* gt_yaw_rate;        /* [S,Tg,A,1]  agent/yaw_rate */
  const uint8_t* tf_mask;          /* [S,Tg,A]    TeacherForcing.get (utils/teacher_forcing.py:33-74) */
  const uint8_t* agent_type;       /* [S,A,3]     */
  const float* agent_size;         /* [S,A,3]     */
  const float* map_boundary;       /* [S,4]  xmin,xmax,ymin,ymax */
  const uint8_t* map_valid;        /* [S,P,20]    raw map (destination polyline nodes, traffic_rule_checker.py:82-98) */
  const uint8_t* map_type;         /* [S,P,11]    */
  const float* map_pos;            /* [S,P,20,2]  */
  const float* map_dir;            /* [S,P,20,2]  */
  const float* goal_gt;            /* [S,A,4] agent/goal, or NULL (test mode: no goal_reached check) */
  /* per scene-mode (B = S*K) */
  const float* latent_sample;      /* [B,A,16]   sampled once per rollout (traffic_bots.py:196-199) */
  const float* latent_logp;        /* [B,A]      */
  const int64_t* dest;             /* [B,A]      destination polyline index */
  const uint8_t* goal_valid;       /* [B,A]      */
  /* tensor-core K|V blocks and key counts from TbSceneOut (may be NULL: the decode step then runs on the CUDA cores) */
  const uint8_t* kv_map_tc;
  const uint8_t* kv_tl_tc;
  const int32_t* n_key_map;
  const int32_t* n_key_tl;
} TbRolloutIn;

typedef struct TbRolloutOut { /* == RolloutBuffer after finish() (utils/buffer.py:72-90), T = n_step */
  float* preds;                    /* [B,A,T,4]  x,y,yaw,spd (pre-override prediction) */
  uint8_t* valid;                  /* [B,A,T]    */
  uint8_t* override_masks;         /* [B,A,T]    */
  float* diffbar_rewards;          /* [B,A,T]    */
  uint8_t* diffbar_rewards_valid;  /* [B,A,T]    */
  float* action_log_probs;         /* [B,A,T]    */
  float* latent_log_probs;         /* [B,A,T]    */
  uint8_t* violations;             /* [6,B,A,T]  order: see TB_N_VIOLATION */
  float* trace_policy_feature;     /* optional [B,A,T,128] (NULL to skip): input of the action head */
  float* trace_action_mean;        /* optional [B,A,T,2] */
} TbRolloutOut;

/* The simulation state that the reference keeps on `Dynamics`, `TrafficBots.hidden`, `TrafficRuleChecker` and
 * `goal_valid` between decode steps lives in one caller-owned buffer.  Field offsets (bytes) for host-side
 * views: */
typedef enum tb_state_field {
  TB_STATE_AGENT_STATE = 0, /* float [B,A,4]  dynamics.agent_state                 */
  TB_STATE_VALID = 1,       /* u8 [2,B,A]     dynamics.agent_valid, double-buffered by step parity (t&1) */
  TB_STATE_KILLED = 2,      /* u8 [B,A]       */
  TB_STATE_VEL = 3,         /* float [B,A,2]  dynamics.vel (only changed by overrides) */
  TB_STATE_ACC = 4,         /* float [B,A]    */
  TB_STATE_YAW_RATE = 5,    /* float [B,A]    */
  TB_STATE_GOAL_VALID = 6,  /* u8 [B,A]       */
  TB_STATE_STICKY = 7,      /* u8 [3,B,A]     outside_map, goal_reached, dest_reached */
  TB_STATE_HIDDEN = 8,      /* float [3,B*A,128]  TrafficBots.hidden (GRU) */
  TB_STATE_N_FIELD = 9
} tb_state_field;
size_t tb_rollout_state_bytes(const TbDims* dims);     /* includes private scratch behind the fields */
size_t tb_rollout_state_offset(const TbDims* dims, int32_t field);

/* rollout set-up: Dynamics.init with frame 0 (waymo_motion.py:251-259), TrafficBots.init (traffic_bots.py:153-161),
 * TrafficRuleChecker.__init__ (traffic_rule_checker.py:45-98), get_goal_feature (goal_manager.py:83-139), and the
 * loop-invariant `mlp_in` halves of add_goal / add_latent (add_latent_goal.py:57). */
int32_t tb_rollout_init(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state, void* stream);

/* decode steps t_first..t_last (1-based, inclusive; WaymoMotion.rollout's loop body, waymo_motion.py:269-343,
 * with WaymoMotion.forward :108-203).  Step t writes slot t-1 of the outputs. */
int32_t tb_rollout_steps(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                         const TbRolloutOut* out, int32_t t_first, int32_t t_last, void* stream);

/* The two halves of one decode step, separately launchable (tb_rollout_steps(t,t) == front(t); back(t)):
 *   front: get_agent_attr_and_pe + agent_encoder + transformer_as2pl + transformer_as2tl (waymo_motion.py:140-155,
 *          traffic_bots.py:205-219) and the K|V projection of the agent<->agent layers;
 *   back:  agent_interaction, agent_temporal, add_goal, add_latent, action head, Dynamics.update/override_states,
 *          TrafficRuleChecker.check, Dynamics.kill, disable_goal_reached, DifferentiableReward.get, buffer write
 *          (traffic_bots.py:228-241, waymo_motion.py:171-203,311-343). */
int32_t tb_step_front(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state, int32_t t,
                      void* stream);
int32_t tb_step_back(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                     const TbRolloutOut* out, int32_t t, void* stream);

/* tb_rollout_init + tb_rollout_steps(1..n_step) == WaymoMotion.rollout (waymo_motion.py:205-354). */
int32_t tb_rollout(const TbDims* dims, const TbRolloutIn* in, const float* packed, void* state,
                   const TbRolloutOut* out, void* stream);

/* ---------------------------------------------------------------- pre-rollout heads ----- */

typedef enum tb_gru {
  TB_GRU_POLICY = 0,        /* model.agent_temporal.rnn                                   */
  TB_GRU_LATENT_PRIOR = 1,  /* model.latent_encoder.agent_temporal_prior.rnn              */
  TB_GRU_LATENT_POST = 2,   /* model.latent_encoder.agent_temporal_post.rnn               */
  TB_GRU_DEST = 3           /* model.goal_manager.goal_predictor.gru_as.rnn               */
} tb_gru;

/* `MultiAgentGRULoop.forward`, 3-D branch (models/modules/agent_temporal.py:133-146: 3-layer GRU over the frames, hidden
 * state of invalid agents zeroed after every frame, outputs zeroed where invalid) fused with the temporal aggregation
 * that follows it in the reference:
 *   mode 0: `TemporalAggregate` max_valid (agent_temporal.py:31-32,43-44)            -> latent encoder (latent_encoder.py:131-137)
 *   mode 1: last valid frame of (GRU output + input)  (goal_manager.py:298-300, last_valid)  -> destination predictor
 * x [n_batch, n_frame, n_agent, 128], valid [n_batch, n_frame, n_agent]; every t_stride-th frame is used.
 * out [n_batch, n_agent, 128], out_valid [n_batch, n_agent] = valid.any(frames).
 * workspace: tb_gru_workspace_bytes (hidden state of the tensor-core kernel), or NULL for the fp32 row-tile kernel. */
size_t tb_gru_workspace_bytes(int32_t n_batch, int32_t n_agent);
int32_t tb_gru_sequence(int32_t which, int32_t mode, const float* x, const uint8_t* valid, int32_t n_batch, int32_t n_frame,
                        int32_t n_agent, int32_t t_stride, const float* packed, void* workspace, float* out, uint8_t* out_valid,
                        void* stream);

typedef enum tb_mlp {
  TB_MLP_LATENT_PRIOR_MEAN = 0, /* model.latent_encoder.latent_prior_dist.mlp_mean : 128 -> 128 -> 16 */
  TB_MLP_LATENT_POST_MEAN = 1   /* model.latent_encoder.latent_post_dist.mlp_mean                     */
} tb_mlp;

/* `MLP.forward` Linear-ReLU-Linear with the valid mask (models/modules/mlp.py:66-85; latent mean: latent_encoder.py:195-199).
 * x [n_row,128], valid [n_row] -> y [n_row, n_out]. */
int32_t tb_mlp_head(int32_t which, const float* x, const uint8_t* valid, int64_t n_row, const float* packed, float* y,
                    void* stream);

/* `DestPredictor.forward`, mode mlp (models/goal_manager.py:228-246,301-307,328-333) after the GRU (tb_gru_sequence
 * TB_GRU_DEST mode 1 gives `tgt`, `tgt_valid`): logits over the P polylines of every agent from the pairwise MLP
 * [map_feature[p], tgt[a]] -> 128 -> 128 -> 1 (LayerNorm + ReLU), polyline-type masks, and the normalisation of
 * `Categorical(logits=)` (models/modules/distributions.py:161-165).
 * map_feature [S,P,128], map_feature_valid [S,P], map_type [S,P,11], tgt [S,A,128], tgt_valid [S,A], agent_type [S,A,3]
 * -> logp [S,A,P], probs [S,A,P].  workspace: tb_dest_workspace_bytes, 256-byte aligned. */
size_t tb_dest_workspace_bytes(int32_t n_scene, int32_t n_agent, int32_t n_pl);
int32_t tb_dest_logits(int32_t n_scene, int32_t n_agent, int32_t n_pl, const float* map_feature,
                       const uint8_t* map_feature_valid, const uint8_t* map_type, const float* tgt, const uint8_t* tgt_valid,
                       const uint8_t* agent_type, const float* packed, void* workspace, float* logp, float* probs, void* stream);

/* ---------------------------------------------------------------- optional rule checks ---- */

/* The four OPTIONAL checks of `TrafficRuleChecker.check` (utils/traffic_rule_checker.py:122-335,420-472: collided,
 * run_road_edge, run_red_light, passive; `enable_check_*` in configs/model/traffic_bots.yaml:240-244) and the collision
 * term of `DifferentiableReward.get` (utils/rewards.py:49-115, `w_collision` > 0).  None of them feeds back into the
 * simulation (only outside_map kills, only goal / dest_reached disable goals), so they are evaluated AFTER the rollout, in
 * parallel over all (scene-mode, step) pairs, from the rollout's outputs: the post-override state the reference checks at
 * step t is `override_masks & ~killed ? GT[t] : preds[t]` with `killed` = running OR of
 * `outside_map_this_step & ~gt_valid` (dynamics.py:132-167).  Results equal the reference's per-step evaluation. */
typedef struct TbRuleIn {
  /* outputs of tb_rollout for the same dims */
  const float* preds;                    /* [B,A,T,4] */
  const uint8_t* valid;                  /* [B,A,T]   */
  const uint8_t* override_masks;         /* [B,A,T]   */
  const uint8_t* outside_map_this_step;  /* [B,A,T] = violations[1] */
  /* ground truth used for overriding (as in TbRolloutIn) */
  const uint8_t* gt_valid;               /* [S,Tg,A]   */
  const float* gt_pos;                   /* [S,Tg,A,2] */
  const float* gt_yaw;                   /* [S,Tg,A,1] */
  const float* gt_spd;                   /* [S,Tg,A,1] */
  const uint8_t* agent_type;             /* [S,A,3] */
  const float* agent_size;               /* [S,A,3] */
  const uint8_t* map_valid;              /* [S,P,20] */
  const uint8_t* map_type;               /* [S,P,11] */
  const float* map_pos;                  /* [S,P,20,2] */
  const float* map_dir;                  /* [S,P,20,2] */
  /* traffic-light stop points handed to the rule checker: the full episode in reactive_replay (waymo_motion.py:439-441),
   * the history frames in joint_future_pred (:523-525); step t reads frame min(t, n_tl_frame - 1) */
  const uint8_t* tl_valid;               /* [S,n_tl_frame,TL]   */
  const float* tl_pos;                   /* [S,n_tl_frame,TL,2] */
  const uint8_t* tl_state;               /* [S,n_tl_frame,TL,5] */
  int32_t n_tl_frame;
  int32_t enable_mask;                   /* bit 0 collided, 1 run_road_edge, 2 run_red_light, 3 passive (needs bit 2: in the
                                            reference `enable_check_passive` alone is a NameError, :441-442,:457-464) */
  float collision_size_scale;            /* 1.1 (traffic_rule_checker.py:28) */
  float w_collision;                     /* differentiable_reward.w_collision; 0 = off */
  int32_t reduce_collision_with_max;     /* differentiable_reward.reduce_collsion_with_max */
} TbRuleIn;

typedef struct TbRuleOut {
  uint8_t* violations;                   /* [8][B,A,T]: collided, collided_this_step, run_road_edge, run_road_edge_this_step,
                                            run_red_light, run_red_light_this_step, passive, passive_this_step */
  float* diffbar_rewards;                /* [B,A,T] in/out (may be NULL when w_collision == 0): the collision term is added to
                                            the imitation reward written by tb_rollout */
  const uint8_t* diffbar_rewards_valid;  /* [B,A,T] (tb_rollout output) */
} TbRuleOut;

size_t tb_rule_workspace_bytes(const TbDims* dims);
int32_t tb_rule_checks(const TbDims* dims, const TbRuleIn* in, const TbRuleOut* out, void* workspace, void* stream);

/* ---------------------------------------------------------------- post-processing + WOMD packing ---- */

/* `WaymoPostProcessing` hyper-parameters (configs/model/traffic_bots.yaml:179-186).  `aggr_thresh` (k-means aggregation) is
 * not implemented: the Python mirror raises for a non-empty list. */
typedef struct TbPostCfg {
  int32_t k_pred;            /* 6 */
  float score_temperature;   /* 1e2; <= 0: off */
  int32_t use_ade;           /* distance between modes: average over the future steps (1) or final displacement (0) */
  int32_t n_mtr;             /* 0 or 3 */
  float mtr_nms_thresh[3];   /* metres, by agent type veh / ped / cyc */
  int32_t n_mpa;             /* 0 or 3 */
  float mpa_nms_thresh[3];
} TbPostCfg;

/* `WaymoPostProcessing.forward` (data_modules/waymo_post_processing.py:33-81) incl. `mtr_nms` (:126-171), `traj_topk`
 * (:173-193; returned in descending score order, the reference's order is unspecified) and `mpa_nms` (:83-124).
 * trajs: element (s, a, mode, t) = 4 floats (x, y, yaw, spd) at trajs + s*stride_scene + a*stride_agent + mode*stride_mode + 4t
 * (strides in floats) -- so the rollout's own [S*K, A, T, 4] output is read in place (stride_scene = K*A*T*4, stride_mode =
 * A*T*4, stride_agent = T*4, trajs = preds + 4*step_future_start) as well as a dense [S,A,n_pred,Tf,4] tensor.
 * scores [S,A,n_pred] (not normalised), valid [S,A], agent_type [S,A,3];  k = min(k_pred, n_pred).
 * -> waymo_trajs [S,Tf,A,k,2], waymo_yaw / waymo_spd [S,Tf,A,k,1] (may be NULL), waymo_scores [S,A,k],
 *    mode_idx [S,A,k] int32 (may be NULL): the input mode each output slot was taken from. */
int32_t tb_post_process(int32_t n_scene, int32_t n_agent, int32_t n_pred, int32_t n_step, const float* trajs,
                        int64_t stride_scene, int64_t stride_agent, int64_t stride_mode, const float* scores,
                        const uint8_t* valid, const uint8_t* agent_type, const TbPostCfg* cfg, float* waymo_trajs,
                        float* waymo_yaw, float* waymo_spd, float* waymo_scores, int32_t* mode_idx, void* stream);

/* `WOMDMetrics.update` (models/metrics/womd.py:60-145, interactive_challenge = False): the per-scene Python loop (:124-138)
 * becomes one CTA per scene.  Every scene's six tensors are written into ONE fixed-size record
 *   [prediction_trajectory [m_joint,K,1,n_ds,2] f32 | prediction_score [m_joint,K] f32 | ground_truth_trajectory [A,step_gt+1,7] f32 |
 *    object_type [A] f32 | ground_truth_is_valid [A,step_gt+1] u8 | prediction_ground_truth_indices_mask [m_joint,1] u8]
 * so that the metrics reduction over the GPUs is a single all-gather of [n_scene, record] bytes (replaces torchmetrics' six
 * per-state gathers, womd.py:23,44-49).  tb_womd_record_bytes returns the record size and the byte offset of each of the six
 * sections, in the order of the reference's states: prediction_trajectory, prediction_score, ground_truth_trajectory,
 * ground_truth_is_valid, prediction_ground_truth_indices_mask, object_type. */
typedef struct TbWomdIn {
  const uint8_t* agent_role;   /* [S,A,3]; [...,2] = agent to predict */
  const uint8_t* agent_valid;  /* [S,n_step_gt_frames,A] */
  const float* agent_pos;      /* [S,n_step_gt_frames,A,2] */
  const float* agent_size;     /* [S,A,3] */
  const float* agent_yaw;      /* [S,n_step_gt_frames,A,1] */
  const float* agent_vel;      /* [S,n_step_gt_frames,A,2] */
  const uint8_t* agent_type;   /* [S,A,3] */
  const float* waymo_trajs;    /* [S,n_step_future,A,K,2] (tb_post_process output) */
  const float* waymo_scores;   /* [S,A,K] or NULL (= uniform 1/K, womd.py:105-106) */
  int32_t n_agent, n_pred, n_step_future, n_step_gt_frames;
  int32_t step_gt, step_current; /* 90, 10 */
  int32_t m_joint;             /* 8 (womd.py:41) */
} TbWomdIn;

typedef struct TbWomdOut {
  float* prediction_trajectory;  /* section pointers of scene 0; scene s lives scene_stride_bytes * s further */
  float* prediction_score;
  float* ground_truth_trajectory;
  uint8_t* ground_truth_is_valid;
  uint8_t* prediction_ground_truth_indices_mask;
  float* object_type;
  int64_t scene_stride_bytes;    /* = tb_womd_record_bytes(...) */
  int32_t* overflow;             /* optional device counter: scenes with more than m_joint agents to predict (the reference
                                    fails with a shape error there); the first m_joint are kept */
} TbWomdOut;

size_t tb_womd_record_bytes(int32_t n_agent, int32_t n_pred, int32_t step_gt, int32_t step_current, int32_t m_joint,
                            int64_t* offsets6);
int32_t tb_womd_pack(int32_t n_scene, const TbWomdIn* in, const TbWomdOut* out, void* stream);

/* Self-test of the tensor-core GEMM machinery (tcgen05.mma, TMEM, bulk-async weight staging, bf16x3 operand split):
 * d[128,128] = a[128,128] @ W^T for packed tensor-core weight block `block` (0 <= block < tb_tc_block_count());
 * mode 0: A operand staged in shared memory, mode 1: A operand in tensor memory. */
int32_t tb_tc_block_count(void);
int32_t tb_tc_first_block(int32_t weight_index); /* -1 if weight i has no tensor-core copy (N or K not multiple of 128) */
int32_t tb_tc_selftest(const float* a, int32_t block, const float* packed, float* d, int32_t mode, void* stream);


/* ------------------------------------------------------------------------------------------------------------------
 * Training primitives (BASELINE.json configs[3]; csrc/tb_train.cu).  The reference trains through torch.autograd
 * (src/pl_modules/waymo_motion.py:356-418 `training_step` + Lightning backward, loss: src/models/metrics/training.py:62-158,
 * optimizer: waymo_motion.py:955-973).  Each entry point is the forward or the hand-derived backward kernel of one
 * differentiable operation of that step; buffers are dense row-major fp32 [rows, cols], masks uint8; "accumulated" outputs
 * are added to, "zero-initialised by the caller" outputs receive atomic partial sums.  The host side that replays them in
 * reverse order is trafficbots_b200/train/{tape,graph}.py.
 *   linear      : nn.Linear (+ReLU) of models/modules/mlp.py:36-64 and the in/out projections of attention.py:85-87,142
 *   layernorm   : nn.LayerNorm(128), eps 1e-5 (+ReLU)
 *   attention   : softmax(QK^T masked, / sqrt(32)) V per head, dead rows (attention.py:89-141,144-146)
 *   gru_gates   : nn.GRU cell gate math (agent_temporal.py:119-153)
 *   masked_max  : map_encoder.py:95-97,105-106 / TemporalAggregate max_valid (agent_temporal.py:31-44)
 *   pair_add, dest_nll : DestPredictor mlp mode (goal_manager.py:294-307,328-333) + goal NLL (metrics/training.py:138-147)
 *   rsample, kl : DiagGaussian.rsample (distributions.py:19-50), BalancedKL with free nats (metrics/loss.py:74-77)
 *   dynamics, reward, sim_flags : utils/dynamics.py:74-167,187-228; utils/rewards.py:117-131;
 *                 utils/traffic_rule_checker.py:101-119,364-410 + models/goal_manager.py:155-161
 *   dropout (drop_seed / drop_site / drop_p arguments, tb_tr_dropout): nn.Dropout in training mode as a counter-based hash mask of
 *                 (per-step seed on the device, site id of the call, element index), regenerated in the backward kernels
 *   sq_norm, adam_step : torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the flat parameter buffer
 * ------------------------------------------------------------------------------------------------------------------ */
/* y = (dropout(relu(x W^T + bias) * keep_lin[row]) + res) * keep_out[row]; bias / keep_lin / res / keep_out / drop_seed may be NULL */
int32_t tb_tr_linear_fwd(const float* x, int64_t M, int32_t K, const float* w, int64_t ldw, int32_t N, const float* bias, int32_t relu, const uint8_t* keep_lin, const float* res, const uint8_t* keep_out, float* y, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
/* dx = dY' W; dw += dY'^T x; db += colsum(dY') with dY' = dy * relu'(y) * dropout mask * rm1[row] * rm2[row] (row masks may be */
/* NULL: the keep_lin / keep_out of the forward; drop_* as in the forward).  dx / dw / db may be NULL (skipped); db needs dw. */
int32_t tb_tr_linear_bwd(const float* dy, const float* x, const float* w, int64_t ldw, const float* y, int32_t relu, const uint8_t* rm1, const uint8_t* rm2, int64_t M, int32_t K, int32_t N, float* dx, float* dw, int64_t lddw, float* db, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
int32_t tb_tr_layernorm_fwd(const float* x, const float* w, const float* b, int32_t relu, int64_t M, int32_t D, float* y, float* stats, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
int32_t tb_tr_layernorm_bwd(const float* dy, const float* x, const float* w, const float* stats, const float* y, int32_t relu, int64_t M, int32_t D, float* dx, float* dw, float* db, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
/* alive[b, s] = 0 for rows without any admissible key (their o and p are 0: attention.py:101-107,144-146), else 1 */
int32_t tb_tr_attention_fwd(const float* q, const float* kv, const uint8_t* key_valid, int32_t eye, int32_t B, int32_t S, int32_t T, float* o, float* p, uint8_t* alive, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
/* dq must be zero-initialised by the caller (partials are added atomically); dkv is overwritten -- or, with kv_batch > 0 (K|V of */
/* batch element b = those of b % kv_batch; kv / dkv hold kv_batch elements), accumulated atomically into a zero-initialised buffer */
int32_t tb_tr_attention_bwd(const float* dout, const float* q, const float* kv, const float* p, const float* o, int32_t B, int32_t S, int32_t T, int32_t kv_batch, float* dq, float* dkv, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
/* y = (a * keep_a[row] + b) * keep[row]; keep_a / b / keep may be NULL */
int32_t tb_tr_add_mask(const float* a, const uint8_t* keep_a, const float* b, const uint8_t* keep, int64_t M, int32_t N, float* y, void* stream);
/* y = x * dropout factor (inter-layer dropout of nn.GRU, agent_temporal.py:116); applied to dy it is its own backward */
int32_t tb_tr_dropout(const float* x, int64_t n, float* y, const uint32_t* drop_seed, uint32_t drop_site, float drop_p, int64_t drop_offset, void* stream);
int32_t tb_tr_axpy(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t M, int32_t N, void* stream);
int32_t tb_tr_select_rows(const uint8_t* mask, const float* a, const float* b, int64_t M, int32_t N, float* y, void* stream);
int32_t tb_tr_select_rows_bwd(const uint8_t* mask, const float* dy, int64_t M, int32_t N, float* da, float* db, void* stream);
int32_t tb_tr_cat2(const float* a, int32_t ka, const float* b, int32_t kb, int64_t M, float* y, void* stream);
int32_t tb_tr_cat2_bwd(const float* dy, int32_t ka, int32_t kb, int64_t M, float* da, float* db, void* stream);
int32_t tb_tr_gru_gates_fwd(const float* gi, const float* gh, const float* h, int64_t M, float* hn, void* stream);
int32_t tb_tr_gru_gates_bwd(const float* dhn, const float* gi, const float* gh, const float* h, int64_t M, float* dgi, float* dgh, float* dh, void* stream);
int32_t tb_tr_masked_max_fwd(const float* x, const uint8_t* valid, int64_t O, int32_t R, int64_t I, int32_t D, float fill, float* y, int32_t* idx, void* stream);
int32_t tb_tr_masked_max_bwd(const float* dy, const int32_t* idx, int64_t O, int32_t R, int64_t I, int32_t D, float* dx, void* stream);
int32_t tb_tr_gather_rows(const float* x, const int64_t* idx, int64_t M, int32_t D, float* y, void* stream);
/* dx must be zero-initialised by the caller */
int32_t tb_tr_scatter_add_rows(const float* dy, const int64_t* idx, int64_t M, int32_t D, float* dx, void* stream);
int32_t tb_tr_pair_add(const float* u, const float* v, int32_t S, int32_t P, int32_t A, float* y, void* stream);
/* dv must be zero-initialised by the caller */
int32_t tb_tr_pair_add_bwd(const float* dy, int32_t S, int32_t P, int32_t A, float* du, float* dv, void* stream);
/* nll_sum [1] must be zero-initialised by the caller */
int32_t tb_tr_dest_nll(const float* logits, const uint8_t* pair_ok, const uint8_t* row_valid, const int64_t* gt, const uint8_t* loss_rows, const float* scale, int64_t n_row, int32_t P, float* nll_sum, float* dlogits, void* stream);
int32_t tb_tr_rsample(const float* mean, const float* log_std, const float* eps, int64_t M, int32_t E, float* z, void* stream);
int32_t tb_tr_rsample_bwd(const float* dz, const float* eps, const float* log_std, int64_t M, int32_t E, float* dlog_std, void* stream);
/* kl_sum [1] zero-initialised by the caller; dlq / dlp are accumulated into */
int32_t tb_tr_kl(const float* mq, const float* lq, const float* mp, const float* lp, const uint8_t* valid, float free_nats, const float* scale, int64_t M, int32_t E, float* kl_sum, float* dmq, float* dmp, float* dlq, float* dlp, void* stream);
int32_t tb_tr_pose_pe(const float* xy, const float* yaw, const float* f_xy, int32_t n_xy, const float* f_yaw, int32_t n_yaw, int64_t M, float* pe, void* stream);
int32_t tb_tr_dir_to_yaw(const float* d, int64_t M, float* yaw, void* stream);
int32_t tb_tr_dynamics(const float* state, const float* mean, const uint8_t* a_type, const uint8_t* valid, int64_t M, float* pred, const float* dpred, float* dstate, float* dmean, void* stream);
int32_t tb_tr_reward(const float* pred, const float* gt, const uint8_t* rv, int64_t M, float* r, const float* dr, float* dpred, void* stream);
int32_t tb_tr_sim_flags(const float* state, const uint8_t* valid, const uint8_t* gt_valid, const float* boundary, const float* dest_pos, const float* dest_dir, const uint8_t* dest_valid, const uint8_t* dest_lane, const uint8_t* dest_edge, const uint8_t* killed, const uint8_t* dest_reached, const uint8_t* goal_valid, int32_t B, int32_t A, uint8_t* o_valid, uint8_t* o_killed, uint8_t* o_dest, uint8_t* o_goal, void* stream);
/* out [1] zero-initialised by the caller */
int32_t tb_tr_masked_sum(const float* x, const uint8_t* mask, int64_t n, float* out, void* stream);
int32_t tb_tr_mask_scale(const uint8_t* mask, const float* scale, int64_t n, float* out, void* stream);
int32_t tb_tr_scale(float* x, int64_t n, float alpha, void* stream);
/* out [1] zero-initialised by the caller */
int32_t tb_tr_sq_norm(const float* g, int64_t n, float* out, void* stream);
int32_t tb_tr_adam_step(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_by_group, const int32_t* group_end, int32_t n_group, float beta1, float beta2, float eps, int32_t step, const float* sq_norm, float max_norm, void* stream);

/* Kernels this library launches on a call path, for accounting (`gpu_launches` in bench.py). */
int64_t tb_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TRAFFICBOTS_B200_H_ */
