// main.cc — `roc_gnn`: the stand-alone driver with the reference's CLI
// (gnn.cc:25-112 top_level_task + :114-179 flags, test.sh / example_run.sh):
//
//   roc_gnn -ll:gpu 1 -lr 0.01 -decay 0.0001 -decay-rate 0.97 -dropout 0.5 \
//           -layers 602-256-41 -file dataset/reddit-dgl -e 3000
//
// One process drives one GPU.  For P > 1 launch P processes with
// ROC_RANK / ROC_WORLD_SIZE set and a shared ROC_NCCL_ID_FILE (rank 0 writes the
// 128-byte NCCL id there); `python -m roc_b200.launch` does that.
// Legion's -ll:* flags other than -ll:gpu are accepted and ignored.
#include <unistd.h>
#include <chrono>
#include <cstring>

#include "host_internal.h"

static bool exchange_nccl_id(int rank, const char* path, unsigned char id[128]) {
  if (rank == 0) {
    if (!Runtime::nccl_unique_id(id)) return false;
    std::string tmp = std::string(path) + ".tmp";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return false;
    fwrite(id, 1, 128, f);
    fclose(f);
    return rename(tmp.c_str(), path) == 0;
  }
  for (int i = 0; i < 600; i++) {
    FILE* f = fopen(path, "rb");
    if (f) {
      size_t n = fread(id, 1, 128, f);
      fclose(f);
      if (n == 128) return true;
    }
    usleep(100000);
  }
  return false;
}

int main(int argc, char** argv) {
  Config config;
  parse_input_args(argv, argc, config);
  const char* er = getenv("ROC_RANK");
  const char* ew = getenv("ROC_WORLD_SIZE");
  int rank = er ? atoi(er) : 0;
  int world = ew ? atoi(ew) : 1;
  if (config.numGPUs <= 0) config.numGPUs = world;
  config.numMachines = 1;
  config.totalGPUs = world;
  fprintf(stderr, "        ===== GNN settings =====\n");
  fprintf(stderr,
          "        dataset = %s seed = %d\n        num_epochs = %d learning_rate = %.4lf\n"
          "        weight_decay = %.4lf dropout_rate = %.4lf\n        decay_rate = %.4lf decay_steps = %d\n",
          config.filename.c_str(), config.seed, config.numEpochs, config.learning_rate, config.weight_decay,
          config.dropout_rate, config.decay_rate, config.decay_steps);
  std::srand(config.seed);   // gnn.cc:56
  fprintf(stderr, "        Layers:");
  for (size_t i = 0; i < config.layers.size(); i++) fprintf(stderr, " %d", config.layers[i]);
  fprintf(stderr, "\n");
  if (config.layers.size() < 2 || config.filename.empty()) {
    fprintf(stderr, "usage: roc_gnn -file <prefix> -layers a-b-c [-e N -lr x -decay x -dropout x ...]\n");
    return 2;
  }
  const char* ed = getenv("ROC_DEVICE");
  Runtime rt(ed ? atoi(ed) : rank, rank, world);
  Runtime* runtime = &rt;
  Context ctx = rt.context();
  if (world > 1) {
    const char* path = getenv("ROC_NCCL_ID_FILE");
    unsigned char id[128];
    if (!path || !exchange_nccl_id(rank, path, id) || !rt.init_nccl(id)) {
      fprintf(stderr, "roc_gnn: NCCL bootstrap failed (set ROC_NCCL_ID_FILE)\n");
      return 3;
    }
  }

  // ---- the model script: same sequence of builder calls as gnn.cc:65-111 ----
  Graph graph(ctx, runtime, config);
  Model model(graph, ctx, runtime);
  const size_t L = config.layers.size();
  Tensor input = model.create_node_tensor<DATATYPE>(config.layers[0]);
  Tensor label = model.create_node_tensor<DATATYPE>(config.layers[L - 1]);
  Tensor mask = model.create_node_tensor<int>(1);
  model.load_features(input, config.filename);
  model.load_labels(label, config.filename);
  model.load_train_mask(mask, config.filename);
  Tensor t = input;
  for (size_t i = 1; i < L; i++) {
    t = model.dropout(t, config.dropout_rate);
    Tensor skip = t;
    t = model.linear(t, config.layers[i], AC_MODE_NONE);
    t = model.indegree_norm(t);
    t = model.scatter_gather(t);
    t = model.indegree_norm(t);
    if (i != L - 1) t = model.relu(t);
    if (L > 3) {   // residual branch, gnn.cc:86-90
      skip = model.linear(skip, (int)t.dims[0], AC_MODE_NONE);
      t = model.add(t, skip);
    }
  }
  model.softmax_cross_entropy(t, label, mask);
  AdamOptimizer* optimizer = new AdamOptimizer(&model, config.learning_rate);
  optimizer->set_weight_decay(config.weight_decay);
  model.optimizer = optimizer;
  model.init(config);

  rt.synchronize();
  auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < config.numEpochs; i++) {
    if ((i != 0) && (i % config.decay_steps == 0)) optimizer->alpha *= config.decay_rate;
    model.train_mode();
    model.zero_gradients();
    model.forward();
    model.backward();
    model.update();
    if (i % 5 == 0) {   // gnn.cc:107-110
      model.infer_mode();
      model.forward();
    }
  }
  rt.synchronize();
  double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (rank == 0)
    fprintf(stderr, "[roc_b200] %d epochs in %.3f s  (%.3f ms/epoch, %.3f M edges/s incl. eval passes)\n",
            config.numEpochs, sec, 1e3 * sec / config.numEpochs,
            1e-6 * (double)graph.numEdges * config.numEpochs / sec);
  roc_sg_plan_destroy(model.myGraph.plan);
  return 0;
}
