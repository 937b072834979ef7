// service_b200_main.cc -- gabby's own InferenceService and HTTP stack with the B200 generator injected through the
// constructor its tests use (/root/reference/src/service.cc:126-129), driven the way /root/reference/src/service_test.cc:28-57
// drives it: POST /v1/chat/completions through http::PostJson, check the envelope, print the assistant's content.
//
//   service_b200 MODEL_DIR SYSTEM_TEXT USER_TEXT [N_REQUESTS]
// prints one line per request on stdout: OBJECT ROLE HEX(content) (hex: the content is arbitrary bytes of a byte-level
// tokenizer); exit code 0 iff every response was a chat.completion.
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>
#include <thread>

#include "b200_generator.h"
#include "http/test_client.h"
#include "http/types.h"
#include "json/json.h"
#include "json/parser.h"
#include "service.h"
#include "utils/logging.h"

int main(int argc, char** argv) {
    using namespace gabby;
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s MODEL_DIR SYSTEM_TEXT USER_TEXT [N_REQUESTS]\n", argv[0]);
        return 2;
    }
    const int n_requests = argc > 4 ? std::atoi(argv[4]) : 1;
    try {
        http::ServerConfig cfg{
            .port = 0,
            .read_timeout_millis = 5'000,
            .write_timeout_millis = 10'000,
            .worker_threads = 2,
        };
        InferenceService service(std::make_unique<http::HttpServer>(cfg), inference::B200Llama3Generator::Load(argv[1], 0, 256, 16));
        service.Start();
        int bad = 0;
        for (int i = 0; i < n_requests; i++) {
            json::ValuePtr request = json::Value::Object({
                {"model", json::Value::String("gabby-1")},
                {"messages", json::Value::Array({
                                 json::Value::Object({{"role", json::Value::String("system")}, {"content", json::Value::String(argv[2])}}),
                                 json::Value::Object({{"role", json::Value::String("user")}, {"content", json::Value::String(argv[3])}}),
                             })},
            });
            json::ValuePtr response = http::PostJson(service.port(), "/v1/chat/completions", request);
            auto obj = response->as_object();
            const std::string object = *obj.at("object")->as_string();
            auto message = obj.at("choices")->as_array()[0]->as_object().at("message")->as_object();
            const std::string content = *message.at("content")->as_string();
            std::string hex;
            for (unsigned char ch : content) {
                char b[3];
                std::snprintf(b, sizeof(b), "%02x", ch);
                hex += b;
            }
            std::cout << "RESPONSE " << object << " " << *message.at("role")->as_string() << " " << hex << std::endl;
            if (object != "chat.completion") bad++;
        }
        service.Stop();
        service.Wait();
        return bad == 0 ? 0 : 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "service_b200: %s\n", e.what());
        return 3;
    }
}
