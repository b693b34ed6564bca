// lstm_main.cc — drop-in host program for the reference's `lstm` binary (R/lstm.cc:50-361), written
// against the C ABI only (include/lstm_b200.h).  With no arguments it reproduces the reference's
// compile-time defaults and its stdout contract:
//   N = 64, M = 256, S = 3, B = 1, lr = 0.1, 1000 epochs, "alice29.txt"        (R/lstm.cc:53-63)
//   "Read <n> bytes (<file>)" / "Empty file! (...)" / "fopen error: (...)"     (:398,410,416)
//   every 100 iterations  "<pct>%\r"  (fixed, width 7, precision 2)            (:274-279)
//   epoch line with t, est GFLOP/s = S(4NM*2+4NN*2)*length / 2^30 / t and avg loss = epoch_loss/(S*length)  (:140,284-291)
//   1000 sampled characters between "************ Generated text |" markers    (:293-356)
// Everything the reference fixes at compile time is a flag here (the reference has no argv).
#include <dirent.h>
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <random>
#include <string>
#include <vector>

#include "../../include/lstm_b200.h"

// rawread (R/lstm.cc:382-420): whole file as bytes; prints exactly what the reference prints
static std::vector<unsigned char> rawread(const char* filename) {
  std::vector<unsigned char> v;
  if (FILE* fp = fopen(filename, "rb")) {
    char buf[1024];
    while (size_t len = fread(buf, 1, sizeof(buf), fp)) v.insert(v.end(), buf, buf + len);
    fclose(fp);
    if (v.size() > 0) std::cout << "Read " << v.size() << " bytes (" << filename << ")" << std::endl;
    else std::cout << "Empty file! (" << filename << ")" << std::endl;
  } else {
    std::cout << "fopen error: (" << filename << ")" << std::endl;
  }
  return v;
}

#define CK(call)                                                                      \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != 0) {                                                                   \
      std::cerr << #call << " failed (" << rc_ << "): " << lstm_last_error(ctx) << std::endl; \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char** argv) {
  size_t N = 64, M = 256, S = 3, B = 1;
  float learning_rate = 1e-1f;
  size_t epochs = 1000;
  std::string load_prefix, save_prefix;
  std::vector<std::string> files;            // default "alice29.txt" (:63); several --file flags / a directory = one corpus
  uint64_t seed = std::random_device{}();   // the reference seeds from std::random_device (:370)
  int dtype = LSTM_F32, stride = 1, device = 0;
  size_t characters_to_generate = 1000;
  float forget_bias = 0.f, state_std = 0.1f;
  long max_iters = -1;
  // held-out evaluation of the last snapshots (OV/lstm_eigen_class_CUDA/lstm.cc:73-86,188-238): off by default
  size_t train_percent = 100;
  double test_every_seconds = 0;
  // the reference's self-test (OV/lstm_eigen_class_batch/lstm.cc:286-318): numerical vs analytic gradients, then exit
  bool gradcheck = false;
  // additions of the last snapshot (OV/lstm_eigen_class_CUDA/lstm.cc): lr = 0 while the window warms up (:364-367: the first
  // 50*S iterations there), the "[Epoch e/E] ... (eta h m s) loss = ... GFlOP/s" progress line (:239-268), a NaN loss is kept
  // out of the running sums (cu_lstm.h:203-215); and of the class_batch snapshot: last-timestep ln loss report, global-max
  // softmax shift (lstm.cc:308-319, lstm.h:175).  --clip is the north-star's fused gradient clipping (0 = off = the reference).
  long lr_warmup = 0;
  bool eta_progress = false;
  double clip = 0.0;
  int loss_mode = 0, softmax_shift = 0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto next = [&]() -> const char* { return (i + 1 < argc) ? argv[++i] : ""; };
    if (a == "--file") files.push_back(next());
    else if (a == "--hidden") N = atol(next());
    else if (a == "--seq") S = atol(next());
    else if (a == "--batch") B = atol(next());
    else if (a == "--epochs") epochs = atol(next());
    else if (a == "--lr") learning_rate = (float)atof(next());
    else if (a == "--seed") seed = strtoull(next(), nullptr, 10);
    else if (a == "--stride") stride = atoi(next());
    else if (a == "--device") device = atoi(next());
    else if (a == "--bf16") dtype = LSTM_BF16;
    else if (a == "--sample") characters_to_generate = atol(next());
    else if (a == "--forget-bias") forget_bias = (float)atof(next());
    else if (a == "--state-std") state_std = (float)atof(next());
    else if (a == "--max-iters") max_iters = atol(next());
    else if (a == "--train-percent") train_percent = atol(next());
    else if (a == "--test-every") test_every_seconds = atof(next());
    else if (a == "--load") load_prefix = next();
    else if (a == "--save") save_prefix = next();
    else if (a == "--gradcheck") gradcheck = true;
    else if (a == "--lr-warmup") lr_warmup = atol(next());
    else if (a == "--progress") eta_progress = std::string(next()) == "eta";
    else if (a == "--clip") clip = atof(next());
    else if (a == "--loss") loss_mode = std::string(next()) == "last-ln" ? 1 : 0;
    else if (a == "--softmax-shift") softmax_shift = std::string(next()) == "global" ? 1 : 0;
    else {
      std::cerr << "usage: lstm [--file F|DIR]... [--hidden N] [--seq S] [--batch B] [--epochs E] [--lr LR] [--seed K] [--stride K]\n"
                   "            [--bf16] [--device D] [--sample N] [--forget-bias X] [--state-std X] [--max-iters K]\n"
                   "            [--load PREFIX] [--save PREFIX] [--train-percent P --test-every SECONDS] [--gradcheck]\n"
                   "            [--lr-warmup ITERS] [--progress percent|eta] [--clip X] [--loss log2-all|last-ln] [--softmax-shift none|global]\n";
      return 2;
    }
  }

  // Multi-corpus input (the reference ships whole corpora next to its snapshots: OV/lstm_eigen_BLAS/cantrbry/, calgary/,
  // reuters21578/ ...): every --file is read with rawread() in the order given, a directory contributes its regular files
  // in name order, and the bytes are concatenated into one training text.
  if (files.empty()) files.push_back("alice29.txt");
  std::vector<unsigned char> data;
  for (const std::string& f : files) {
    std::vector<std::string> members;
    if (DIR* d = opendir(f.c_str())) {
      while (struct dirent* e = readdir(d)) {
        const std::string path = f + "/" + e->d_name;
        struct stat sb;
        if (stat(path.c_str(), &sb) == 0 && S_ISREG(sb.st_mode)) members.push_back(path);
      }
      closedir(d);
      std::sort(members.begin(), members.end());
    } else {
      members.push_back(f);
    }
    for (const std::string& m : members) {
      const std::vector<unsigned char> part = rawread(m.c_str());
      data.insert(data.end(), part.begin(), part.end());
    }
  }
  if (data.empty()) return 0;   // the reference carries on with a 0x0 matrix and does nothing

  lstm_ctx* ctx = nullptr;
  int rc = lstm_create(&ctx, (int)M, (int)N, (int)S, (int)B, device, dtype);
  if (rc != 0) {
    std::cerr << "lstm_create failed (" << rc << "): " << lstm_last_error(nullptr) << std::endl;
    return 1;
  }
  CK(lstm_init_params(ctx, seed, 0.01f, forget_bias));              // :113-119
  if (clip > 0) CK(lstm_set_option(ctx, LSTM_OPT_CLIP, clip));
  if (loss_mode) CK(lstm_set_option(ctx, LSTM_OPT_LOSS_MODE, loss_mode));
  if (softmax_shift) CK(lstm_set_option(ctx, LSTM_OPT_SOFTMAX_SHIFT, softmax_shift));
  if (!load_prefix.empty() && lstm_load_text_ckpt(ctx, load_prefix.c_str()) != 0)
    std::cout << lstm_last_error(ctx) << std::endl;                  // the reference prints and keeps the random init
  // first train_percent % for training, the rest held out (class_CUDA/lstm.cc:73-86)
  std::vector<unsigned char> testdata;
  if (train_percent < 100) {
    const size_t percent_size = data.size() / 100;
    testdata.assign(data.begin() + train_percent * percent_size, data.end());
    data.resize(train_percent * percent_size);
    std::cout << "Train set size: " << data.size() << ", Test set size: " << testdata.size() << ", Total: "
              << data.size() + testdata.size() << std::endl;
  }
  CK(lstm_load_text(ctx, data.data(), data.size()));
  const size_t length = data.size();
  auto last_test = std::chrono::steady_clock::now();
  double window_loss = 0.0;
  size_t window_iters = 0, results_rows = 0;
  if (B > 1) {                                                        // OV/lstm_eigen_opt/lstm.cc:140-144
    std::mt19937_64 g(seed + 17);
    std::vector<uint64_t> pos(B);
    for (auto& p : pos) p = g() % (length - S) + S;
    CK(lstm_set_positions(ctx, pos.data()));
  }
  if (gradcheck) {
    // window of every stream taken from the text at its start position: x_t = data[p - S + t], target_t = data[p - S + t + 1]
    std::vector<uint64_t> pos(B);
    CK(lstm_get_positions(ctx, pos.data()));
    std::vector<int32_t> xw(S * B, -1), tw(S * B, -1);
    for (size_t b = 0; b < B; b++)
      for (size_t t = 1; t < S; t++) {
        const size_t p = (size_t)pos[b] - S + t;
        xw[t * B + b] = data[p % length];
        tw[t * B + b] = data[(p + 1) % length];
      }
    CK(lstm_reset_state(ctx, seed + 3, state_std));
    double rep[5][6];
    int passed = 0;
    CK(lstm_gradcheck(ctx, xw.data(), tw.data(), 100, seed, 1e-5, rep, nullptr, nullptr, nullptr, &passed));
    const char* names[5] = {"W", "U", "b", "Why", "by"};
    const int order[5] = {LSTM_U, LSTM_W, LSTM_WHY, LSTM_B, LSTM_BY};       // check_gradients(): U, W, Why, b, by
    for (int q = 0; q < 5; q++) {                                            // layout of check_gradient_error()
      const int w = order[q];
      std::cout << std::endl << std::setw(15) << std::setprecision(12) << "[" << names[w] << "]" << std::endl
                << std::setw(20) << " numerical range (" << std::setw(20) << rep[w][2] << ", " << std::setw(20) << rep[w][3] << ")" << std::endl
                << std::setw(20) << " analytical range (" << std::setw(20) << rep[w][4] << ", " << std::setw(20) << rep[w][5] << ")" << std::endl
                << std::setw(20) << " max rel. error " << std::setw(20) << rep[w][0];
      if (rep[w][0] > 1e-1) std::cout << std::setw(23) << "!!!  >1e-1 !!!";
      std::cout << std::endl << std::setw(20) << " mean rel. error " << std::setw(20) << rep[w][1];
      if (rep[w][1] > 1e-3) std::cout << std::setw(23) << "!!!  >1e-3  !!!";
      std::cout << std::endl;
    }
    std::cout << std::endl << (passed ? "gradient check OK" : "gradient check FAILED") << std::endl;
    lstm_destroy(ctx);
    return passed ? 0 : 1;
  }
  const double flops_per_epoch = (double)S * (4.0 * N * M * 2 + 4.0 * N * N * 2) * (double)length;   // :140
  const size_t iters_per_epoch = (length - S + stride - 1) / stride;
  std::vector<double> losses(100);
  long total_iters = 0, nan_iters = 0;
  const double flops_per_iteration = (double)(S - 1) * (double)B * (22.0 * N * M + 24.0 * N * N);

  // Seeding follows the ORDER in which the reference constructs its generators, the k-th one seeded seed + k
  // (the test build of the reference program, tests/golden/make_ref_run.py, is given the same schedule): W, U, Why
  // (k = 0, 1, 2, inside lstm_init_params), then per epoch h, c (:146-147), _h, _c (:306-307) and the sampling
  // generator (:309-310) — so `lstm --seed K` and the reference started with REF_SEED=K walk the same trajectory
  // up to float rounding.
  uint64_t gen_k = seed + 3;
  for (size_t e = 0; e < epochs; e++) {
    double epoch_loss = 0.0;
    if (B == 1 && state_std != 0.f) {
      // randn(h,0,0.1); randn(c,0,0.1) draw the whole N x S matrices in (row, col) order (:146-147, :375); the column
      // that the first shift (:163-164) moves into slot 0 is column 1.
      std::vector<float> h0(N), c0(N);
      std::mt19937 mh((uint32_t)gen_k), mc((uint32_t)(gen_k + 1));
      std::normal_distribution<> dh(0.0f, state_std), dc(0.0f, state_std);
      for (size_t i = 0; i < N; i++)
        for (size_t j = 0; j < S; j++) {
          const float vh = (float)dh(mh), vc = (float)dc(mc);
          if (j == 1) { h0[i] = vh; c0[i] = vc; }
        }
      CK(lstm_set_state(ctx, h0.data(), c0.data()));
    } else {
      CK(lstm_reset_state(ctx, gen_k, state_std));                    // N x B draws for the carried-in column
    }
    gen_k += 2;
    auto t0 = std::chrono::steady_clock::now();
    size_t done = 0;
    bool stop = false;
    while (done < iters_per_epoch && !stop) {
      // the reference prints when i % 100 == 0 (i = S + iteration index): run up to the next multiple of 100
      const size_t i_now = S + done * stride;
      size_t chunk = (100 - (i_now % 100) + stride - 1) / stride;
      if (chunk == 0) chunk = 100 / stride ? 100 / stride : 1;
      if (chunk > iters_per_epoch - done) chunk = iters_per_epoch - done;
      if (max_iters >= 0 && total_iters + (long)chunk >= max_iters) { chunk = (size_t)(max_iters - total_iters); stop = true; }
      if (total_iters < lr_warmup && total_iters + (long)chunk > lr_warmup) { chunk = (size_t)(lr_warmup - total_iters); stop = false; }
      const float lr_now = total_iters < lr_warmup ? 0.f : learning_rate;   // class_CUDA/lstm.cc:364-367
      if (chunk > losses.size()) losses.resize(chunk);
      const auto chunk_t0 = std::chrono::steady_clock::now();
      if (chunk > 0) CK(lstm_train_text(ctx, (int)chunk, stride, lr_now, losses.data()));
      for (size_t k = 0; k < chunk; k++) {
        if (std::isnan(losses[k])) { nan_iters++; continue; }          // cu_lstm.h:210: a NaN never reaches the running sums
        epoch_loss += losses[k]; window_loss += losses[k];
      }
      window_iters += chunk;
      done += chunk;
      if (test_every_seconds > 0 && testdata.size() > 1) {
        const auto now = std::chrono::steady_clock::now();
        const double since = std::chrono::duration<double>(now - last_test).count();
        if (since >= test_every_seconds) {
          double test_bpc = 0;
          CK(lstm_eval_bpc(ctx, testdata.data(), testdata.size(), &test_bpc));
          const double train_bpc = window_loss / ((double)window_iters * (double)(S - 1));
          const double gflops = (double)window_iters * (double)(S - 1) * (double)B * (22.0 * N * M + 24.0 * N * N) / since / 1073741824.0;
          std::cout << std::endl << "Train error: " << train_bpc << ", Test error: " << test_bpc << std::endl;
          // the results row: idx, secs since last, train, test, GFlOP/s (class_CUDA/lstm.cc:205-226), printed and logged
          char row[160];
          snprintf(row, sizeof(row), "%zu %g %g %g %g", results_rows, since, train_bpc, test_bpc, gflops);
          std::cout << row << std::endl << std::endl;
          if (!save_prefix.empty()) {
            FILE* rf = fopen((save_prefix + ".txt").c_str(), "a");
            if (rf) { fprintf(rf, "%s\n", row); fclose(rf); }
            CK(lstm_save_text_ckpt(ctx, save_prefix.c_str()));
            // 5000 sampled characters next to the checkpoint (class_CUDA/lstm.cc:228-234; h = c = 0 there: reset_std = 0)
            std::vector<uint8_t> smp(5000);
            CK(lstm_sample(ctx, seed + 7919 * (results_rows + 1), nullptr, nullptr, smp.data(), smp.size(), 0));
            FILE* sf = fopen((save_prefix + "_sample.txt").c_str(), "wb");
            if (sf) { fwrite(smp.data(), 1, smp.size(), sf); fclose(sf); }
          }
          results_rows++;
          window_loss = 0.0; window_iters = 0;
          last_test = std::chrono::steady_clock::now();
        }
      }
      total_iters += (long)chunk;
      const size_t i = S + done * stride;
      if (eta_progress && chunk > 0) {                  // class_CUDA/lstm.cc:239-268
        const double chunk_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - chunk_t0).count();
        size_t eta_sec = (size_t)(chunk_time / (double)chunk * (double)(iters_per_epoch - done));
        const size_t eta_hours = eta_sec / 3600, eta_min = (eta_sec % 3600) / 60;
        eta_sec = eta_sec % 60;
        const double gflops_per_sec = (double)chunk * flops_per_iteration / 1073741824.0 / chunk_time;
        std::cout << std::setw(15) << "[Epoch " << e + 1 << "/" << epochs << "]" << std::fixed << std::setw(10) << std::setprecision(2)
                  << 100.0f * (float)i / (float)length << "%     (eta " << std::setw(2) << eta_hours << " h " << std::setfill('0')
                  << std::setw(2) << eta_min << " m " << std::setfill('0') << std::setw(2) << eta_sec << " s)" << std::setfill(' ')
                  << std::setw(12) << std::setprecision(6) << "loss = " << epoch_loss / ((double)done * (double)(S - 1))
                  << std::setw(9) << std::setprecision(2) << gflops_per_sec << " GFlOP/s\r" << std::flush;
      } else if ((i % 100 == 0 && i < length) || stride > 1)   // the reference's loop ends at i = length - 1 (:151): no print at i = length
        std::cout << std::fixed << std::setw(7) << std::setprecision(2) << 100.0f * (float)i / (float)length << "%\r" << std::flush;
    }
    CK(lstm_sync(ctx));
    const double epoch_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::cout << std::endl
              << "====================================================================================" << std::endl
              << "Epoch " << e + 1 << "/" << epochs << std::fixed << std::setprecision(3) << ", t = " << epoch_time << " s"
              << ", est GFLOP/s = " << (flops_per_epoch / powf(2, 30)) / epoch_time << ", avg loss = "
              << epoch_loss / ((double)S * (double)length) << " bits/char" << std::endl;

    // sampling (:293-356): _h, _c ~ N(0, 0.1), then 1000 draws
    std::vector<float> h0(N), c0(N);
    {
      std::mt19937 mh((uint32_t)gen_k), mc((uint32_t)(gen_k + 1));
      std::normal_distribution<> dh(0.0f, 0.1f), dc(0.0f, 0.1f);
      for (auto& v : h0) v = (float)dh(mh);
      for (auto& v : c0) v = (float)dc(mc);
    }
    std::vector<uint8_t> text(characters_to_generate);
    CK(lstm_sample(ctx, gen_k + 2, h0.data(), c0.data(), text.data(), text.size(), 0));
    gen_k += 3;
    std::cout << std::endl << std::endl << "************ Generated text |";
    for (uint8_t ch : text) std::cout << (char)ch;
    std::cout << "| Generated text END ************" << std::endl;
    if (nan_iters) std::cout << nan_iters << " iteration(s) with a NaN loss were left out of the averages" << std::endl;
    if (!save_prefix.empty()) CK(lstm_save_text_ckpt(ctx, save_prefix.c_str()));
    if (stop) break;
  }
  lstm_destroy(ctx);
  return 0;
}
